"""Similarity-transform fit (Umeyama + RANSAC) of the inference pose estimate, batched.

SURVEY.md section 8f-2: the step right after the hot path at test time.  The reference (model/util/umeyama.py:9-41,
95-201) runs, per image, a Python loop of 100 RANSAC rounds -- each a 5-point closed-form fit (one 3x3 SVD), a residual
over every correspondence and several host synchronisations.  Here all rounds of all images are one batched fit (one
batched SVD) and one batched residual evaluation; the sequential rules of the reference loop are applied to the residual
table afterwards, so results AND the consumption of the global CPU random generator are those of the reference:

  * round i samples `torch.randint(0, n, (5,))` from the global CPU generator (:104);
  * a round replaces the incumbent only when its residual is strictly smaller (:107), the incumbent starts at 1e10;
  * the loop ends after the first round that leaves the incumbent below the stop threshold (:111-112) -- later rounds
    draw no random numbers;
  * the final transform is the closed-form fit over the inliers of the winning round (:33).

Closed form (:165-201, restated; X source, Y target, n points): cov = (Y-mean Y)(X-mean X)^T / n, U D V^T = svd(cov) with
the last column of U and the last singular value negated when det U det V^T < 0, R = (U V^T)^T, s = sum(D) / sum of the
unbiased per-axis variances of X, t = mean Y - mean X . (s R)   [row vector times matrix], transform = [s R | t].
Quirk kept: the residual applies s R to column vectors while t is built with the row-vector product (the transpose).

The public functions keep the reference's names and argument meaning; `fit_similarity_batch` is the batched entry the
pose fit (model/pose_fit.py) uses.  Pure torch, device-agnostic (CPU for the parity tests against the reference module).
"""
import torch

N_ROUNDS = 100      # umeyama.py:22
N_SAMPLE = 5        # :104
_START_RESIDUAL = 1e10


def _det3(m):
    """Determinant of (..., 3, 3) matrices by cofactor expansion (only its sign is used; avoids batched LU launches)."""
    return (m[..., 0, 0] * (m[..., 1, 1] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 1])
            - m[..., 0, 1] * (m[..., 1, 0] * m[..., 2, 2] - m[..., 1, 2] * m[..., 2, 0])
            + m[..., 0, 2] * (m[..., 1, 0] * m[..., 2, 1] - m[..., 1, 1] * m[..., 2, 0]))


def _closed_form(src, tgt, weight=None, strict=True):
    """src, tgt (..., n, 3) rows; weight (..., n) bool or None (all points) -> s (...), R (..., 3, 3), t (..., 3), bad (...).
    bad marks problems whose covariance holds a NaN: with `strict` that raises like the reference (:180-184), otherwise
    those rows come back as NaN."""
    if weight is None:
        n = src.shape[-2]
        cs, ct = src.mean(-2), tgt.mean(-2)
        xs, xt = src - cs[..., None, :], tgt - ct[..., None, :]
        cov = xt.transpose(-1, -2) @ xs / n
        var = xs.pow(2).sum((-1, -2)) / (n - 1)
    else:
        w = weight[..., None]
        zero = torch.zeros((), device=src.device, dtype=src.dtype)
        n = weight.sum(-1).to(src.dtype)
        cs, ct = torch.where(w, src, zero).sum(-2) / n[..., None], torch.where(w, tgt, zero).sum(-2) / n[..., None]
        xs, xt = torch.where(w, src - cs[..., None, :], zero), torch.where(w, tgt - ct[..., None, :], zero)
        cov = xt.transpose(-1, -2) @ xs / n[..., None, None]
        var = xs.pow(2).sum((-1, -2)) / (n - 1)
    return _from_moments(cs, ct, cov, var, strict)


def _from_moments(cs, ct, cov, var, strict=True):
    """Closed form from the means cs / ct (..., 3), the covariance cov (..., 3, 3) of the centred points (target rows x source
    columns, divided by n) and the unbiased source variance var (...)."""
    bad = torch.isnan(cov).any(-1).any(-1)
    if strict:
        if bad.any():
            raise RuntimeError('There are NANs in the input.')
    else:       # LAPACK / cuSOLVER refuse non-finite input: decompose a stand-in and blank the result below
        cov = torch.where(bad[..., None, None], torch.eye(3, device=cov.device, dtype=cov.dtype), cov)
    U, D, Vh = torch.linalg.svd(cov, full_matrices=True)
    sign = torch.where(_det3(U) * _det3(Vh) < 0.0, -1.0, 1.0).to(cov.dtype)
    D = torch.cat((D[..., :2], D[..., 2:] * sign[..., None]), -1)
    U = torch.cat((U[..., :, :2], U[..., :, 2:] * sign[..., None, None]), -1)
    R = (U @ Vh).transpose(-1, -2)
    s = D.sum(-1) / var
    if not strict:
        s = torch.where(bad, torch.full_like(s, float('nan')), s)
    t = ct - (cs[..., None, :] @ (s[..., None, None] * R))[..., 0, :]
    return s, R, t, bad


def _point_residuals(s, R, t, src, tgt):
    """|tgt - (s R src + t)| per point: s (..., h), R (..., h, 3, 3), t (..., h, 3), src/tgt (..., n, 3) -> (..., h, n)."""
    A = s[..., None, None] * R
    pred = torch.einsum('...hij,...nj->...hni', A, src) + t[..., None, :]
    return torch.linalg.norm(tgt[..., None, :, :] - pred, dim=-1)


def _sequential_choice(residual, stop_t, bad):
    """The reference loop replayed over a residual table (..., h).  Returns (winning round, last executed round, found):
    `found` False when no round was accepted or the loop hit a round whose sample had a NaN covariance (`bad`; the
    reference raises there, so that round is the last one executed)."""
    h = residual.shape[-1]
    inf = torch.full_like(residual, float('inf'))
    r = torch.where(residual < _START_RESIDUAL, residual, inf)          # NaN or >= 1e10: never accepted (:107)
    ar = torch.arange(h, device=residual.device)
    first = lambda flag: torch.where(flag, ar, torch.full_like(ar, h)).amin(-1)
    stop_at, raise_at = first(torch.cummin(r, -1).values < stop_t[..., None]), first(bad)
    last = torch.minimum(torch.minimum(stop_at, raise_at), torch.full_like(stop_at, h - 1))
    r = torch.where(ar <= last[..., None], r, inf)
    best = r.argmin(-1)                                                  # first occurrence = strict '<' of the loop
    found = torch.isfinite(r.gather(-1, best[..., None])[..., 0]) & (raise_at > last)
    return best, last, found


def _thresholds(src, tgt, valid=None):
    """Pass / stop thresholds of estimateSimilarityTransform (:16-21): the larger of the two mean-norm ratios, /100."""
    ns, nt = torch.linalg.norm(src, dim=-1), torch.linalg.norm(tgt, dim=-1)
    if valid is None:
        ms, mt = ns.mean(-1), nt.mean(-1)
    else:
        cnt = valid.sum(-1)
        ms, mt = torch.where(valid, ns, 0 * ns).sum(-1) / cnt, torch.where(valid, nt, 0 * nt).sum(-1) / cnt
    ts, st = mt / ms, ms / mt
    pass_t = torch.where(st > ts, st, ts)
    return pass_t, pass_t / 100


def fit_similarity_batch(src, tgt, counts, max_table_bytes=1 << 30):
    """RANSAC + Umeyama for B images at once.

    src, tgt: (B, n_max, 3) correspondences, image b's counts[b] real ones first (the rest is padding, any finite
    value); counts: python ints.  Returns (scale (B,), rotation (B,3,3), translation (B,3), ok (B,) bool on the host
    side as a list).  ok[b] False = the reference would not have produced a transform for image b (no correspondence,
    no accepted round, or fewer than 10 % inliers, :29-31) -- scale / rotation / translation rows are then undefined.
    Random numbers: drawn from the global CPU generator exactly as B consecutive reference calls would (see module
    docstring); when some image ends its loop early the generator is rewound and the later images are re-drawn."""
    B, n_max = src.shape[0], src.shape[1]
    dev, dt = src.device, src.dtype
    scale = torch.ones(B, device=dev, dtype=dt)
    rot = torch.eye(3, device=dev, dtype=dt).repeat(B, 1, 1)
    trans = torch.zeros(B, 3, device=dev, dtype=dt)
    ok = [False] * B
    start = 0
    while start < B:
        live = [b for b in range(start, B) if counts[b] > 0]      # n = 0: randint raises in the reference, nothing drawn
        if not live:
            break
        states, draws = {}, []
        for b in live:
            states[b] = torch.get_rng_state()
            draws.append(torch.randint(0, counts[b], (N_ROUNDS, N_SAMPLE)))
        idx = torch.stack(draws).to(dev)                                              # L, H, 5
        sel = torch.tensor(live, device=dev)
        s_l, t_l = src.index_select(0, sel), tgt.index_select(0, sel)
        cnt = torch.tensor([counts[b] for b in live], device=dev)
        valid = torch.arange(n_max, device=dev)[None] < cnt[:, None]                  # L, n
        pass_t, stop_t = _thresholds(s_l, t_l, valid)
        rows = torch.arange(len(live), device=dev)[:, None, None]
        hs, hR, ht, hbad = _closed_form(s_l[rows, idx], t_l[rows, idx], strict=False)  # (L,H), (L,H,3,3), (L,H,3), (L,H)
        native = src.is_cuda and dt == torch.float32        # the two point-cloud passes as kernels (csrc/scp_posefit.cu)
        if native:
            from ...ops import posefit as K
            cnt32 = cnt.to(torch.int32)
            residual = K.residual_table(s_l, t_l, cnt32, hs, hR, ht)
        else:
            # residual table, a slab of rounds at a time
            slab = max(1, min(N_ROUNDS, max_table_bytes // max(1, len(live) * n_max * 3 * src.element_size() * 2)))
            residual = torch.empty(len(live), N_ROUNDS, device=dev, dtype=dt)
            for h0 in range(0, N_ROUNDS, slab):
                pr = _point_residuals(hs[:, h0:h0 + slab], hR[:, h0:h0 + slab], ht[:, h0:h0 + slab], s_l, t_l)
                residual[:, h0:h0 + slab] = torch.linalg.norm(torch.where(valid[:, None], pr, 0 * pr), dim=-1)
        best, last, found = _sequential_choice(residual, stop_t, hbad)
        pick = lambda x: x[torch.arange(len(live), device=dev), best]
        if native:
            m = K.inlier_moments(s_l, t_l, cnt32, pick(hs), pick(hR), pick(ht), pass_t, found)
            ratio_ok = m['n_inliers'] / cnt.to(dt) >= 0.1
            fs, fR, ft, _ = _from_moments(m['mean_src'], m['mean_tgt'], m['cov'] / m['n_used'][:, None, None],
                                          m['sq'] / (m['n_used'] - 1), strict=False)
        else:
            pr = _point_residuals(pick(hs)[:, None], pick(hR)[:, None], pick(ht)[:, None], s_l, t_l)[:, 0]
            inlier = (pr < pass_t[:, None]) & valid
            ratio_ok = inlier.sum(-1).to(dt) / cnt.to(dt) >= 0.1
            safe = inlier | ~(found & ratio_ok)[:, None] & valid        # keep the closed form finite for rejected images
            fs, fR, ft, _ = _closed_form(s_l, t_l, safe, strict=False)
        last_h, ok_h = last.tolist(), (found & ratio_ok).tolist()                     # the one host synchronisation
        early = next((i for i, l in enumerate(last_h) if l < N_ROUNDS - 1), None)
        done = len(live) if early is None else early + 1
        sel_done = sel[:done]
        scale[sel_done], rot[sel_done], trans[sel_done] = fs[:done], fR[:done], ft[:done]
        for i in range(done):
            ok[live[i]] = bool(ok_h[i])
        if early is None:
            break
        # image live[early] left its loop after round last_h[early]: put the generator where the reference leaves it
        b = live[early]
        torch.set_rng_state(states[b])
        torch.randint(0, counts[b], (last_h[early] + 1, N_SAMPLE))
        start = b + 1
    return scale, rot, trans, ok


# ---- the reference's single-problem interface (same names, argument meaning and return values) ----------------------

def estimateSimilarityUmeyama(SourceHom, TargetHom):
    """(4, n) homogeneous columns -> (Scales (3,), Rotation (3,3), Translation (1,3), OutTransform (4,4))."""
    src, tgt = SourceHom[:3].transpose(0, 1), TargetHom[:3].transpose(0, 1)
    s, R, t, _ = _closed_form(src, tgt)
    out = torch.eye(4, device=src.device, dtype=src.dtype)
    out[:3, :3] = s * R
    out[:3, 3] = t
    return s.reshape(-1).repeat(3), R, t[None], out


def evaluateModel(OutTransform, SourceHom, TargetHom, PassThreshold):
    """(residual norm over all points, inlier ratio, inlier indices)."""
    per_point = torch.linalg.norm((TargetHom - OutTransform @ SourceHom)[:3], dim=0)
    inliers = (per_point < PassThreshold).nonzero().reshape(-1)
    return torch.linalg.norm(per_point), inliers.shape[0] / SourceHom.shape[1], inliers


def getRANSACInliers(SourceHom, TargetHom, MaxIterations=100, PassThreshold=200, StopThreshold=1):
    """Inlier columns of the winning round and their ratio; all rounds evaluated at once."""
    n = SourceHom.shape[1]
    dev, dt = SourceHom.device, SourceHom.dtype
    src, tgt = SourceHom[:3].transpose(0, 1), TargetHom[:3].transpose(0, 1)
    state = torch.get_rng_state()
    idx = torch.randint(0, n, (MaxIterations, N_SAMPLE)).to(dev)
    hs, hR, ht, hbad = _closed_form(src[idx], tgt[idx], strict=False)
    residual = torch.linalg.norm(_point_residuals(hs, hR, ht, src, tgt), dim=-1)
    stop_t = torch.as_tensor(StopThreshold, device=dev, dtype=dt)
    best, last, found = _sequential_choice(residual, stop_t, hbad)
    if int(last) < MaxIterations - 1:           # rounds after the last executed one draw nothing in the reference
        torch.set_rng_state(state)
        torch.randint(0, n, (int(last) + 1, N_SAMPLE))
    if bool(hbad[int(last)]):
        raise RuntimeError('There are NANs in the input.')
    if not bool(found):                         # no round ever accepted: the reference indexes with a float arange and raises
        raise IndexError('no RANSAC round was accepted (all residuals non-finite or >= 1e10)')
    per_point = _point_residuals(hs[best][None], hR[best][None], ht[best][None], src, tgt)[0]
    inliers = (per_point < PassThreshold).nonzero().reshape(-1)
    return SourceHom[:, inliers], TargetHom[:, inliers], inliers.shape[0] / n


def estimateSimilarityTransform(source, target, verbose=False):
    """source, target (n, 3) -> (Scales, Rotation, Translation, OutTransform), or four Nones below 10 % inliers."""
    ones = torch.ones((source.shape[0], 1), device=source.device, dtype=source.dtype)
    SourceHom, TargetHom = torch.cat([source, ones], 1).transpose(0, 1), torch.cat([target, ones], 1).transpose(0, 1)
    pass_t, stop_t = _thresholds(source, target)
    if verbose:
        print('Pass threshold: ', pass_t)
        print('Stop threshold: ', stop_t)
        print('Number of iterations: ', N_ROUNDS)
    src_in, tgt_in, ratio = getRANSACInliers(SourceHom, TargetHom, MaxIterations=N_ROUNDS, PassThreshold=pass_t,
                                             StopThreshold=stop_t)
    if ratio < 0.1:
        print('[ WARN ] - Something is wrong. Small BestInlierRatio: ', ratio)
        return None, None, None, None
    out = estimateSimilarityUmeyama(src_in, tgt_in)
    if verbose:
        print('BestInlierRatio:', ratio)
        print('Rotation:\n', out[1])
        print('Translation:\n', out[2])
        print('Scales:', out[0])
    return out


def estimateRestrictedAffineTransform(source, target, verbose=False):
    raise NotImplementedError      # as in the reference (:43-44)
