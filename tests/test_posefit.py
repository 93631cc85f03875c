"""Inference pose fit: batched RANSAC + Umeyama and `PoseFitter.pose_fitting` against outputs of
the REFERENCE (model/util/umeyama.py, model/tester.py:324-427) recorded in tests/golden/posefit_golden.npz by
tests/golden/make_posefit_golden.py -- results and the number of random integers consumed from the global generator."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'posefit_golden.npz')


@pytest.fixture(scope='module')
def gold():
    g = np.load(GOLD)
    return lambda k: torch.from_numpy(g[k])


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_single_fit_matches_reference_outputs_and_rng_consumption(gold):
    from self_corr_pose_b200.model.util import umeyama as U
    for i in range(int(gold('fit_cases'))):
        torch.manual_seed(100 + i)
        scales, rot, trans, transform = U.estimateSimilarityTransform(gold('fit%d_src' % i), gold('fit%d_tgt' % i))
        assert torch.equal(torch.randint(0, 1 << 30, (4,)), gold('fit%d_next_draw' % i)), i    # same generator state
        for got, name in ((scales, 'scales'), (rot, 'rotation'), (trans, 'translation'), (transform, 'transform')):
            want = gold('fit%d_%s' % (i, name))
            assert got.shape == want.shape and _rel(got, want) < 1e-5, (i, name)


def test_batched_fit_equals_consecutive_reference_calls(gold):
    """All cases in ONE padded batch, seeded once: image b must see the random integers the b-th consecutive reference
    call would draw, including after the two cases whose loop ends in its first round."""
    from self_corr_pose_b200.model.util import umeyama as U
    n_case = int(gold('fit_cases'))
    probs = [(gold('fit%d_src' % i), gold('fit%d_tgt' % i)) for i in range(n_case)]
    probs.insert(2, (torch.zeros(0, 3), torch.zeros(0, 3)))              # an image without correspondences draws nothing
    counts = [p[0].shape[0] for p in probs]
    S, T = torch.zeros(len(probs), max(counts), 3), torch.zeros(len(probs), max(counts), 3)
    for b, (s, t) in enumerate(probs):
        S[b, :counts[b]], T[b, :counts[b]] = s, t
    torch.manual_seed(77)
    want = [U.estimateSimilarityTransform(s, t) if s.shape[0] else None for s, t in probs]
    state = torch.get_rng_state()
    torch.manual_seed(77)
    scale, rot, trans, ok = U.fit_similarity_batch(S, T, counts)
    assert torch.equal(torch.get_rng_state(), state)
    assert ok == [w is not None for w in want]
    for b, w in enumerate(want):
        if w is not None:
            assert _rel(scale[b], w[0][0]) < 1e-5 and _rel(rot[b], w[1]) < 1e-5 and _rel(trans[b], w[2][0]) < 1e-5, b
    # a small residual-table budget (several slabs of rounds) changes nothing
    torch.manual_seed(77)
    scale2, rot2, trans2, ok2 = U.fit_similarity_batch(S, T, counts, max_table_bytes=1 << 20)
    assert ok2 == ok and torch.allclose(scale2, scale, rtol=1e-6) and torch.allclose(rot2, rot, atol=1e-6)


def test_sequential_rules_of_the_ransac_loop():
    from self_corr_pose_b200.model.util.umeyama import _sequential_choice
    inf, nan = float('inf'), float('nan')
    r = torch.tensor([[5., 3., 3., 0.5, 0.1, 9.],       # stop threshold 1: round 3 ends the loop, round 4 is never run
                      [5., nan, 2., 2., 7., 4.],        # strict '<': the first of two equal residuals wins
                      [nan, 2e10, inf, nan, 1e10, nan], # nothing below the 1e10 start value: no transform
                      [4., 3., 2., 1.5, 9., 9.]])       # a NaN covariance in round 2 raises in the reference
    bad = torch.zeros(4, 6, dtype=torch.bool)
    bad[3, 2] = True
    best, last, found = _sequential_choice(r, torch.tensor([1., 1., 1., 1.]), bad)
    assert best[:2].tolist() == [3, 2] and last.tolist() == [3, 5, 5, 2] and found.tolist() == [True, True, False, False]


def _pose_inputs(gold, dev):
    size = int(gold('pose_size'))
    opts = SimpleNamespace(img_size=size, base_rot=gold('pose_base_rot').tolist())
    T = lambda k: gold('pose_' + k).to(dev)
    B = T('mask').shape[0]
    batch = (torch.zeros(B, 3, size, size, device=dev), T('mask'), T('depth'), None, None, None, None, T('foc'), None,
             T('pp'), None, None)
    pred = (T('pred_v'), None, None, None, T('match'), T('conf'))
    return opts, batch, pred


def _check_pose(gold, dev, tol):
    from self_corr_pose_b200.model.pose_fit import PoseFitter
    opts, batch, pred = _pose_inputs(gold, dev)
    torch.manual_seed(9)
    out = PoseFitter(opts, device=dev).pose_fitting(batch, pred)
    assert torch.equal(torch.randint(0, 1 << 30, (4,)), gold('pose_next_draw'))
    for got, name in zip(out, ('bbox', 'verts', 'rotation', 'translation')):
        want = gold('pose_' + name)
        assert got.shape == want.shape and _rel(got.cpu(), want) < tol, name
    # images 3 and 4 have no / too few usable pixels: default pose (identity before the base rotation, 0.5 m ahead)
    assert torch.allclose(out[3][3].cpu(), torch.tensor([[0., 0., 0.5]]))


def test_pose_fitting_matches_reference_tester(gold):
    _check_pose(gold, 'cpu', 1e-5)


@pytest.mark.gpu
def test_pose_fitting_on_gpu_matches_reference_tester(gold):
    _check_pose(gold, 'cuda', 1e-4)


def test_eval_match_confidence_matches_reference():
    """Correspondence.match_confidence (evaluation-mode output consumed by the pose fit) against the confidence map the
    reference's Correspondence.match produced in evaluation mode (tests/golden/make_matchconf_golden.py)."""
    from self_corr_pose_b200.model.module.correspondence import Correspondence, make_meshgrid
    g = np.load(os.path.join(os.path.dirname(GOLD), 'matchconf_golden.npz'))
    T = lambda k: torch.from_numpy(g[k])
    hf = 16
    corr = Correspondence.__new__(Correspondence)          # no device state needed for this method
    corr.hf = corr.wf = hf
    corr.meshgrid = make_meshgrid(hf, hf, 'cpu')
    match_up = T('match')                                  # nearest upsampling 16 -> 64: every 4th pixel is the source
    match_lr = match_up[:, :, ::4, ::4].permute(0, 2, 3, 1).reshape(match_up.shape[0], hf * hf, 3)
    conf = corr.match_confidence(match_lr, T('imatch'), T('pred_v'), T('mask'))
    want = T('match_conf')
    assert conf.shape == want.shape and torch.equal(conf == 0, want == 0)
    assert _rel(conf, want) < 1e-5


def test_oracle_fit_pinned_to_reference(gold):
    """oracle/posefit.py (the reference's sequential formulation, used for timing and large-size checks on the GPU box)
    against the recorded reference outputs."""
    from oracle import posefit as O
    for i in range(int(gold('fit_cases'))):
        torch.manual_seed(100 + i)
        s, R, t = O.fit(gold('fit%d_src' % i), gold('fit%d_tgt' % i))
        assert torch.equal(torch.randint(0, 1 << 30, (4,)), gold('fit%d_next_draw' % i)), i
        assert _rel(s, gold('fit%d_scales' % i)[0]) < 1e-5 and _rel(R, gold('fit%d_rotation' % i)) < 1e-5
        assert _rel(t, gold('fit%d_translation' % i)[0]) < 1e-5


def test_oracle_pose_fitting_pinned_to_reference(gold):
    from oracle import posefit as O
    T = lambda k: gold('pose_' + k)
    torch.manual_seed(9)
    out = O.pose_fitting(T('mask'), T('depth'), T('match'), T('conf'), T('foc'), T('pp'), T('pred_v'), T('base_rot'),
                         int(T('size')))
    assert torch.equal(torch.randint(0, 1 << 30, (4,)), T('next_draw'))
    for got, name in zip(out, ('bbox', 'verts', 'rotation', 'translation')):
        assert _rel(got, T(name)) < 1e-5, name


def test_batched_pose_fit_equals_oracle_on_a_larger_batch():
    """Random 64 x 64 evaluation batch (thousands of correspondences per image, beyond the golden's size)."""
    from oracle import posefit as O
    from self_corr_pose_b200.model.pose_fit import PoseFitter
    size, B, N = 64, 4, 200
    g = torch.Generator().manual_seed(4)
    foc = (3.5 + 0.3 * torch.rand(B, 2, generator=g)).double()
    pp = (0.05 * torch.randn(B, 2, generator=g)).double()
    match = torch.rand(B, 3, size, size, generator=g) - 0.5
    A = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    cam = 300 * torch.einsum('bchw,bcd->bdhw', match, A) + torch.tensor([0., 0., 900.])[None, :, None, None]
    depth = (cam[:, 2] + 3 * torch.randn(B, size, size, generator=g)) * (torch.rand(B, size, size, generator=g) > 0.1)
    mask = (torch.rand(B, size, size, generator=g) > 0.3).float()
    conf = torch.rand(B, 1, size, size, generator=g)
    pred_v = torch.rand(B, N, 3, generator=g) - 0.5
    base = torch.eye(3).reshape(-1)
    torch.manual_seed(1)
    want = O.pose_fitting(mask, depth, match, conf, foc, pp, pred_v, base, size)
    state = torch.get_rng_state()
    torch.manual_seed(1)
    fitter = PoseFitter(SimpleNamespace(img_size=size), device='cpu')
    got = fitter.pose_fitting((None, mask, depth, None, None, None, None, foc, None, pp, None, None),
                              (pred_v, None, None, None, match, conf))
    assert torch.equal(torch.get_rng_state(), state)
    for a, b in zip(got, want):
        assert _rel(a, b) < 1e-5


def test_sequential_choice_equals_a_literal_loop_on_random_tables():
    """300 random residual tables (ties, NaN, inf, values over the 1e10 start, NaN-covariance rounds, stop thresholds that
    fire early / late / never) against the reference loop written out literally (umeyama.py:102-112)."""
    from self_corr_pose_b200.model.util.umeyama import _sequential_choice
    g = torch.Generator().manual_seed(0)
    H, n_tab = 12, 300
    r = torch.rand(n_tab, H, generator=g).mul(8).round() / 2                     # coarse values: many exact ties
    special = torch.rand(n_tab, H, generator=g)
    r[special < 0.08] = float('nan')
    r[(special >= 0.08) & (special < 0.12)] = float('inf')
    r[(special >= 0.12) & (special < 0.16)] = 3e10
    bad = torch.rand(n_tab, H, generator=g) < 0.03
    stop = torch.rand(n_tab, generator=g).mul(4).round() / 2                     # 0 (never) .. 2
    best, last, found = _sequential_choice(r, stop, bad)
    for t in range(n_tab):
        inc, inc_i, executed, raised = 1e10, None, H - 1, False
        for i in range(H):
            executed = i
            if bad[t, i]:
                raised = True
                break
            if float(r[t, i]) < inc:
                inc, inc_i = float(r[t, i]), i
            if inc < float(stop[t]):
                break
        assert int(last[t]) == executed, t
        assert bool(found[t]) == (inc_i is not None and not raised), t
        if inc_i is not None and not raised:
            assert int(best[t]) == inc_i, t


@pytest.mark.gpu
def test_posefit_kernels_match_host_formulation():
    """csrc/scp_posefit.cu through the C ABI (ops/posefit.py): the residual table of 100 candidates x 3 images and the inlier
    moments of the winning rounds against the torch statements of model/util/umeyama.py evaluated in fp64 -- including a
    short image (37 points), padding rows, outliers and an image whose fit is rejected (all real points are used)."""
    from self_corr_pose_b200.model.util import umeyama as U
    from self_corr_pose_b200.ops import posefit as K
    L, n, H = 3, 9000, 100
    g = torch.Generator().manual_seed(5)
    src = torch.rand(L, n, 3, generator=g) - 0.5
    R0 = torch.linalg.qr(torch.randn(L, 3, 3, generator=g))[0]
    tgt = 300 * src @ R0 + torch.tensor([0., 0., 900.]) + 3 * torch.randn(L, n, 3, generator=g)
    tgt[:, ::7] += 80 * torch.randn(L, (n + 6) // 7, 3, generator=g)
    counts = torch.tensor([n, n - 4001, 37], dtype=torch.int32)
    valid = torch.arange(n)[None] < counts[:, None]
    idx = torch.stack([torch.randint(0, int(c), (H, 5), generator=g) for c in counts])
    rows = torch.arange(L)[:, None, None]
    hs, hR, ht, _ = U._closed_form(src[rows, idx], tgt[rows, idx], strict=False)
    pr = U._point_residuals(hs.double(), hR.double(), ht.double(), src.double(), tgt.double())
    want = torch.linalg.norm(torch.where(valid[:, None], pr, 0 * pr), dim=-1)
    c = lambda t: t.cuda()
    got = K.residual_table(c(src), c(tgt), c(counts), c(hs), c(hR), c(ht))
    assert _rel(got.cpu(), want) < 1e-5
    best = want.argmin(-1)
    pick = lambda x: x[torch.arange(L), best]
    pass_t, _ = U._thresholds(src, tgt, valid)
    found = torch.tensor([True, False, True])
    prb = U._point_residuals(pick(hs)[:, None], pick(hR)[:, None], pick(ht)[:, None], src, tgt)[:, 0]
    inl = (prb < pass_t[:, None]) & valid
    ok = found & (inl.sum(-1).float() / counts.float() >= 0.1)
    safe = torch.where(ok[:, None], inl, valid)
    fs, fR, ft, _ = U._closed_form(src.double(), tgt.double(), safe, strict=False)
    m = K.inlier_moments(c(src), c(tgt), c(counts), c(pick(hs)), c(pick(hR)), c(pick(ht)), c(pass_t), c(found))
    assert torch.equal(m['n_used'].cpu().long(), safe.sum(-1)) and torch.equal(m['n_inliers'].cpu().long(), inl.sum(-1))
    gs, gR, gt, _ = U._from_moments(m['mean_src'], m['mean_tgt'], m['cov'] / m['n_used'][:, None, None],
                                    m['sq'] / (m['n_used'] - 1), strict=False)
    print('PARITY posefit table %.2e scale %.2e rotation %.2e translation %.2e' %
          (_rel(got.cpu(), want), _rel(gs.cpu(), fs), _rel(gR.cpu(), fR), _rel(gt.cpu(), ft)))
    assert _rel(gs.cpu(), fs) < 1e-5 and _rel(gR.cpu(), fR) < 1e-5 and _rel(gt.cpu(), ft) < 1e-5
    m2 = K.inlier_moments(c(src), c(tgt), c(counts), c(pick(hs)), c(pick(hR)), c(pick(ht)), c(pass_t), c(found))
    assert all(torch.equal(m[k], m2[k]) for k in m)          # fixed-order reductions


@pytest.mark.gpu
def test_batched_fit_on_gpu_equals_host_fit_with_the_same_draws():
    """fit_similarity_batch on CUDA tensors (kernels) against the same call on the CPU (torch statements pinned to the
    reference above): same generator consumption, same verdicts, transforms within fp32 rounding."""
    from self_corr_pose_b200.model.util import umeyama as U
    B, n = 5, 6000
    g = torch.Generator().manual_seed(8)
    src = torch.rand(B, n, 3, generator=g) - 0.5
    R0 = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    tgt = 250 * src @ R0 + torch.tensor([10., -20., 800.]) + 2 * torch.randn(B, n, 3, generator=g)
    tgt[:, ::5] += 60 * torch.randn(B, (n + 4) // 5, 3, generator=g)
    counts = [n, 4500, 0, 800, 5]
    torch.manual_seed(3)
    s0, R0_, t0, ok0 = U.fit_similarity_batch(src, tgt, counts)
    state = torch.get_rng_state()
    torch.manual_seed(3)
    s1, R1, t1, ok1 = U.fit_similarity_batch(src.cuda(), tgt.cuda(), counts)
    assert torch.equal(torch.get_rng_state(), state) and ok0 == ok1
    keep = torch.tensor(ok0)
    assert keep.any()
    assert _rel(s1.cpu()[keep], s0[keep]) < 1e-4 and _rel(R1.cpu()[keep], R0_[keep]) < 1e-4 and _rel(t1.cpu()[keep], t0[keep]) < 1e-4
