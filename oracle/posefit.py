"""CPU restatement (plain torch, the reference's one-image-at-a-time, one-round-at-a-time formulation) of the inference
pose fit -- TEST INFRASTRUCTURE ONLY (tests/ and tools/ timing scripts; never imported by the package).

Follows:
  closed_form()        model/util/umeyama.py:165-201   (estimateSimilarityUmeyama)
  ransac()             model/util/umeyama.py:95-130    (getRANSACInliers + evaluateModel)
  fit()                model/util/umeyama.py:9-41      (estimateSimilarityTransform)
  pose_fitting()       model/tester.py:324-427         (per-image compaction, back-projection, fit, box)
Pinned against the reference's own functions by tests/golden/make_posefit_golden.py -> posefit_golden.npz
(tests/test_posefit.py::test_oracle_*).  Random integers come from torch's global CPU generator, like the reference's.
"""
import torch


def closed_form(X, Y):
    """X, Y (3, n) columns -> (s, R (3,3), t (3,), T (4,4)) with T = [s R | t] (conventions of umeyama.py:188-199)."""
    n = X.shape[1]
    mx, my = X.mean(1), Y.mean(1)
    cov = (Y - my[:, None]) @ (X - mx[:, None]).T / n
    if torch.isnan(cov).any():
        raise RuntimeError('There are NANs in the input.')
    U, D, Vh = torch.linalg.svd(cov)
    if torch.linalg.det(U) * torch.linalg.det(Vh) < 0:
        D, U = D.clone(), U.clone()
        D[2], U[:, 2] = -D[2], -U[:, 2]
    R = (U @ Vh).T
    s = D.sum() / X.var(dim=1).sum()
    t = my - mx @ (s * R)
    T = torch.eye(4, dtype=X.dtype, device=X.device)
    T[:3, :3], T[:3, 3] = s * R, t
    return s, R, t, T


def ransac(X, Y, rounds, pass_t, stop_t):
    """Indices of the inliers of the best of `rounds` 5-point fits (smallest residual norm over ALL points; the loop
    ends once the incumbent residual is under stop_t) and their ratio."""
    n = X.shape[1]
    best, best_inl, ratio = 1e10, None, 0
    for _ in range(rounds):
        pick = torch.randint(0, n, (5,))
        T = closed_form(X[:, pick], Y[:, pick])[3]
        err = torch.linalg.norm(Y - (T[:3, :3] @ X + T[:3, 3:]), dim=0)
        total = torch.linalg.norm(err)
        if total < best:
            best, best_inl = total, (err < pass_t).nonzero().reshape(-1)
            ratio = best_inl.shape[0] / n
        if best < stop_t:
            break
    if best_inl is None:
        raise IndexError('no round accepted')          # the reference indexes with a float arange here and raises
    return best_inl, ratio


def fit(source, target):
    """source, target (n, 3) -> (s, R, t) or None when fewer than 10 % of the points are inliers."""
    X, Y = source.T, target.T
    ts = torch.linalg.norm(target, dim=1).mean() / torch.linalg.norm(source, dim=1).mean()
    pass_t = max(ts, 1 / ts) if not torch.isnan(ts) else ts
    inl, ratio = ransac(X, Y, 100, pass_t, pass_t / 100)
    if ratio < 0.1:
        return None
    return closed_form(X[:, inl], Y[:, inl])[:3]


def pose_fitting(mask, depth, match, match_conf, foc_crop, pp_crop, pred_v, base_rot, size):
    """-> (bbox (B,9,3), posed vertices (B,N,3), rotation (B,3,3), translation (B,1,3))."""
    B = mask.shape[0]
    c = (torch.arange(size, dtype=torch.float32) + 0.5) / (size / 2) - 1
    gx, gy = c[None].expand(size, size).reshape(-1), c[:, None].expand(size, size).reshape(-1)
    keep = ((depth > 0)[:, None] * mask[:, None] * match_conf).reshape(B, -1) > 0
    Rs, ts, ss = [], [], []
    for i in range(B):
        k = keep[i]
        K = torch.eye(3)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2] = foc_crop[i, 0], foc_crop[i, 1], pp_crop[i, 0], pp_crop[i, 1]
        rays = torch.stack((gx[k], gy[k], torch.ones(int(k.sum()))), 1) @ K.inverse().T
        cam = rays * depth[i].reshape(-1)[k][:, None] / rays[:, 2:]
        model = match[i].reshape(3, -1)[:, k].T
        try:
            out = fit(model, cam)
        except (RuntimeError, IndexError):
            out = None
        if out is None:        # the reference only handles the exception; a None result is given the same default here
            out = torch.tensor(100.), torch.eye(3), torch.tensor([0., 0., 500.])
        ss.append(out[0].reshape(1).repeat(3))
        Rs.append(out[1])
        ts.append(out[2])
    rotation, translation = torch.stack(Rs), (torch.stack(ts) * 0.001)[:, None]
    scale = (torch.stack(ss) * 0.001)[:, None]
    base = base_rot.reshape(1, 3, 3).expand(B, -1, -1)
    v = pred_v.bmm(base.transpose(1, 2))
    rotation = base.bmm(rotation)
    lo, hi = v.min(1).values, v.max(1).values
    corners = [(lo + hi) / 2]
    for ix in (lo, hi):
        for iy in (lo, hi):
            for iz in (lo, hi):
                corners.append(torch.stack((ix[:, 0], iy[:, 1], iz[:, 2]), -1))
    bbox = torch.stack(corners, 1)
    return (bbox * scale).bmm(rotation) + translation, (v * scale).bmm(rotation) + translation, rotation, translation
