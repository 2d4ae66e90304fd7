"""CPU model of FAST mode's proof obligations (fast.cuh: hit_fragile, tie_margin).

FAST mode answers with the nearest triangle among ALL triangles the exact test accepts; the
reference answers with the nearest among those its own rounded box tests let it reach.  The device
code hands a ray to the exact traversal whenever it cannot prove the two agree.  This file restates
the exact triangle test, hit_fragile and tie_margin in numpy (same IEEE operations in the same
order) and checks the CLAIM behind them on far more rays than the GPU fuzz can afford: whenever the
brute-force answer differs from the oracle's traversal, the ray is flagged.  (An unflagged
difference would be a hole in the proof; flagged-but-equal rays only cost time.)"""
import numpy as np
import pytest

import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene

K_ROBUST = 1e-13
DET_MIN = 1e-11


def _tri_exact(a, ba, ca, o, d):
    """Triangle::intersection (triangle.rs:49-80) for rays (R,3) x triangles (N,3) -> (R,N) arrays."""
    c0, c1 = ba[None, :, :], ca[None, :, :]
    c2 = -d[:, None, :]
    x, y, z = 0, 1, 2
    det = (c0[..., x] * (c1[..., y] * c2[..., z] - c2[..., y] * c1[..., z]) - c1[..., x] * (c0[..., y] * c2[..., z] - c2[..., y] * c0[..., z])
           + c2[..., x] * (c0[..., y] * c1[..., z] - c1[..., y] * c0[..., z]))

    def cross(p, q):
        return np.stack([(p[..., y] * q[..., z]) - (p[..., z] * q[..., y]), (p[..., z] * q[..., x]) - (p[..., x] * q[..., z]),
                         (p[..., x] * q[..., y]) - (p[..., y] * q[..., x])], -1)

    def dot(p, q):
        return (p[..., x] * q[..., x] + p[..., y] * q[..., y]) + p[..., z] * q[..., z]

    with np.errstate(all="ignore"):
        r0, r1, r2 = cross(c1, c2) / det[..., None], cross(c2, c0) / det[..., None], cross(c0, c1) / det[..., None]
        w = o[:, None, :] - a[None, :, :]
        u, v, t = dot(r0, w), dot(r1, w), dot(r2, w)
        hit = ~(np.abs(det) < DET_MIN) & ~((u < 0.0) | (v < 0.0) | (1.0 < u + v) | (t < 0.0))
    return hit, t


def _hit_fragile(a, ba, ca, o, d, t, bound):
    """fast.cuh hit_fragile for matched arrays (R,3) of the winning triangle per ray."""
    b, c = a + ba, a + ca
    lo, hi = np.minimum(a, np.minimum(b, c)), np.maximum(a, np.maximum(b, c))
    E = K_ROBUST * (bound + np.abs(o))
    D = np.abs(d)
    pos, neg = d > 0.0, d < 0.0
    N = np.where(pos, lo - o, np.where(neg, o - hi, 0.0))
    F = np.where(pos, hi - o, np.where(neg, o - lo, 0.0))
    zero = ~pos & ~neg
    frag = np.any(zero & (~((o - lo > E) & (hi - o > E)) | ~(d == 0.0)), axis=1)
    moving = D > 0.0
    frag |= np.any(moving & ~(F > E), axis=1)
    frag |= np.any(moving & ~(t[:, None] * D - N > -E), axis=1)
    for i in range(3):
        for j in range(3):
            if i != j:
                both = moving[:, i] & moving[:, j]
                frag |= both & ~(F[:, j] * D[:, i] - N[:, i] * D[:, j] > E[:, i] * D[:, j] + E[:, j] * D[:, i])
    return frag


def _tie_margin(o, d, t, bound):
    with np.errstate(all="ignore"):
        m = np.max(np.where(d != 0.0, (bound + np.abs(o)) / np.abs(d), 0.0), axis=1)
    return K_ROBUST * (2.0 * m + np.abs(t))


def _check(orc, tri, rays_o, rays_d):
    raw = RawScene(width=4, height=4)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5)}]
    raw.tri_pos = tri
    raw.tri_mat = np.zeros(len(tri), np.uint32)
    scene = orc.OracleScene(raw)
    n = len(tri)
    rec = np.stack([scene.triangle(T.BVH_TRIANGLES, i)[0] for i in range(n)])      # post-build order
    a, ba, ca = rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]
    bound = float(np.nextafter(np.float32(np.abs(scene.bvh(T.BVH_TRIANGLES)["bounds"][0]).max()), np.float32(np.inf)))
    ref = scene.trace_rays(rays_o, rays_d)[0]
    unflagged_differences = flagged = differences = 0
    for s in range(0, len(rays_o), 2000):
        o, d = rays_o[s:s + 2000], rays_d[s:s + 2000]
        ok = np.isfinite(o).all(1) & np.isfinite(d).all(1)              # NaN rays: FAST answers "miss" directly
        hit, t = _tri_exact(a, ba, ca, o, d)
        tt = np.where(hit, t, np.inf)
        win = np.argmin(tt, axis=1)
        has = hit.any(axis=1)
        tw = tt[np.arange(len(o)), win]
        tt2 = tt.copy()
        tt2[np.arange(len(o)), win] = np.inf
        t2 = tt2.min(axis=1)
        with np.errstate(all="ignore"):
            frag = has & _hit_fragile(a[win], ba[win], ca[win], o, d, tw, bound)
            tie = has & np.isfinite(t2) & ~(t2 - tw > _tie_margin(o, d, t2, bound))
        flag = frag | tie
        r = ref[s:s + 2000]
        r_has = r["kind"] >= 0
        same = (has == r_has) & (~has | ((win == r["prim"]) & (tw.view(np.uint64) == r["t"].view(np.uint64))))
        differences += int((~same & ok).sum())
        unflagged_differences += int((~same & ~flag & ok).sum())
        flagged += int((flag & ok).sum())
    return differences, unflagged_differences, flagged


@pytest.mark.parametrize("flavour", ["far", "scales", "slivers", "coplanar", "grid"])
def test_every_disagreement_with_the_reference_traversal_is_flagged(orc, flavour):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fz", os.path.join(os.path.dirname(__file__), "test_gpu_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    total_diff = total_flag = total = 0
    for seed in range(int(os.environ.get("RT_FUZZ_SEEDS", 4))):
        rng = np.random.default_rng(7000 + 10 * seed + len(flavour))
        tri = fz._soup(rng, int(rng.integers(150, 500)), flavour)
        o, d, _ = fz._rays(rng, tri, 12000)
        diff, unflagged, flagged = _check(orc, tri, o, d)
        assert unflagged == 0, (flavour, seed, diff, unflagged)
        total_diff += diff
        total_flag += flagged
        total += len(o)
    # the inputs are adversarial: the reference's traversal really does disagree with brute force
    assert total_flag > 0
    print(f"{flavour}: {total} rays, {total_diff} differ from brute force, {total_flag} flagged")
