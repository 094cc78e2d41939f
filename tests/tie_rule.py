"""The "documented float tie" rule of the feature-list parity contract, as an executable check shared by
the GPU tests (CUDA lists vs cv2) and the CPU tests (oracle lists vs cv2)."""
import numpy as np

from oracle import image_oracle as io


def as_list(p):
    return np.zeros((0, 2), np.float32) if p is None else np.asarray(p).reshape(-1, 2)


TIE_REL = 2.0 ** -20          # candidates whose lambda_min differ by <= 2^-20 * max lambda_min may swap


def cv2_eig(img, bs):
    """cv2.cornerMinEigenVal where cv2 is importable (the GPU box runs the same image as the build container); the
    scalar oracle's map otherwise (same arithmetic as the GPU: exact integer window sums)."""
    try:
        import cv2
        return cv2.cornerMinEigenVal(img, bs)
    except ImportError:
        return io.min_eig_map(img, bs)


def explain_by_ties(got, ref, lam_map):
    """PROVES the documented-tie claim for one pair of ordered corner lists. Returns the number of tie groups.
    When the lists differ they must hold the same corners, and the permutation between them must decompose into
    blocks (maximal runs of positions over which both lists hold the same corners) inside which every corner's
    lambda_min ON THE REFERENCE'S OWN MAP lies within TIE_REL * max(lambda_min) of every other: only candidates OpenCV
    itself could not tell apart beyond its fp32 summation noise ever trade places."""
    got, ref = as_list(got), as_list(ref)
    if got.shape == ref.shape and np.array_equal(got, ref):
        return 0
    assert got.shape == ref.shape, "corner counts differ: %d vs %d" % (len(got), len(ref))
    key = lambda p: (int(p[0]), int(p[1]))
    gk, rk = [key(p) for p in got], [key(p) for p in ref]
    assert sorted(gk) == sorted(rk), "the two lists do not hold the same corners"
    tol = TIE_REL * float(lam_map.max())
    groups, i, n = 0, 0, len(gk)
    while i < n:
        if gk[i] == rk[i]:
            i += 1
            continue
        j = i + 1
        seen_g, seen_r = {gk[i]}, {rk[i]}
        while seen_g != seen_r:
            assert j < n, "unbalanced permutation block"
            seen_g.add(gk[j]); seen_r.add(rk[j]); j += 1
        lam = np.array([lam_map[y, x] for (x, y) in rk[i:j]], dtype=np.float64)
        assert lam.max() - lam.min() <= tol, ("corners trade places although cv2's lambda_min tells them apart",
                                               rk[i:j], lam.tolist(), tol)
        groups += 1
        i = j
    return groups
