"""ORACLE tooling (test infrastructure, never shipped, never on the product path).

Independent cross-checks for the two scikit-image 0.20.0 internals that oracle/grasp_tail.py restates and that cannot be
executed here (scikit-image is not installable offline).  Everything below is built from *installed third-party code*
(scipy, OpenCV), not from the oracle's own loops, so that a shared misreading cannot pass silently:

* ``ensure_spacing_literal`` / ``peak_local_max_literal`` — the control flow of skimage/_shared/coord.py:ensure_spacing
  and skimage/feature/peak.py as published for 0.20.0, written out literally: the 50 / +100 / +200 … batch split
  (``min_split_size=50``, ``max_split_size=2000``), one ``scipy.spatial.cKDTree.query_ball_point(r=spacing, p=inf)`` per
  batch, ``scipy.spatial.distance.cdist(..., minkowski, p=inf)`` with the strict ``d < spacing`` rejection rule, the
  ``max_out`` early break, ``np.delete`` of the rejected rows and the final truncation; the peak mask comes from the real
  ``scipy.ndimage.maximum_filter(footprint=ones((5,5)), mode='nearest')``.  oracle/grasp_tail.py claims that all of this
  equals ONE greedy pass over the stably sorted candidates; tests/test_oracle_tail.py asserts the claim on plateau-heavy
  maps (thousands of tied candidates, so several batches and the early break are exercised).
* ``polygon_cv2`` — skimage.draw.polygon(r, c, shape) fills every lattice point for which its point-in-polygon routine
  returns inside / vertex / edge.  For integer vertices that set is {p : cv2.pointPolygonTest(quad, p, False) >= 0}
  (OpenCV evaluates the test in exact integer arithmetic for CV_32S contours and integral query points).  Used to verify
  oracle/grasp_tail.py:polygon on >= 10 000 random truncated ``cv2.boxPoints`` quadrilaterals.

What stays unverifiable offline is stated in oracle/grasp_tail.py's header.
"""
from __future__ import annotations

import numpy as np


def _ensure_spacing_batch(coord: np.ndarray, spacing, p_norm, max_out):
    from scipy.spatial import cKDTree, distance

    tree = cKDTree(coord)
    indices = tree.query_ball_point(coord, r=spacing, p=p_norm)
    rejected = set()
    naccepted = 0
    for idx, candidates in enumerate(indices):
        if idx not in rejected:
            candidates = list(candidates)
            candidates.remove(idx)
            if candidates:
                dist = distance.cdist([coord[idx]], coord[candidates], distance.minkowski, p=p_norm).reshape(-1)
                candidates = [c for c, d in zip(candidates, dist) if d < spacing]
            rejected.update(candidates)
            naccepted += 1
            if max_out is not None and naccepted >= max_out:
                break
    output = np.delete(coord, tuple(rejected), axis=0)
    if max_out is not None:
        output = output[:max_out]
    return output


def ensure_spacing_literal(coords: np.ndarray, spacing=1, p_norm=np.inf, min_split_size=50, max_out=None,
                           max_split_size=2000) -> np.ndarray:
    output = coords
    if len(coords):
        coords = np.atleast_2d(coords)
        if min_split_size is None:
            batch_list = [coords]
        else:
            coord_count = len(coords)
            split_idx = [min_split_size]
            split_size = min_split_size
            while coord_count - split_idx[-1] > max_split_size:
                split_size *= 2
                split_idx.append(split_idx[-1] + min(split_size, max_split_size))
            batch_list = np.array_split(coords, split_idx)
        output = np.zeros((0, coords.shape[1]), dtype=coords.dtype)
        for batch in batch_list:
            output = _ensure_spacing_batch(np.vstack([output, batch]), spacing, p_norm, max_out)
            if max_out is not None and len(output) >= max_out:
                break
    return output


def peak_local_max_literal(image: np.ndarray, min_distance: int = 2, threshold_abs: float = 0.4, num_peaks=5) -> np.ndarray:
    """peak_local_max(image, min_distance, threshold_abs, num_peaks) with exclude_border=True, labels=None, p_norm=inf."""
    from scipy import ndimage as ndi

    image = np.asarray(image)
    size = 2 * min_distance + 1
    footprint = np.ones((size,) * image.ndim, dtype=bool)
    image_max = ndi.maximum_filter(image, footprint=footprint, mode="nearest")
    out = image == image_max
    if np.all(out):  # "no peak for a trivial image"
        out[:] = False
    out &= image > threshold_abs
    b = min_distance  # exclude_border=True -> border width min_distance on every axis
    for ax in range(out.ndim):
        sl = [slice(None)] * out.ndim
        sl[ax] = slice(None, b); out[tuple(sl)] = False
        sl[ax] = slice(-b, None); out[tuple(sl)] = False
    coord = np.nonzero(out)
    intensities = image[coord]
    idx_maxsort = np.argsort(-intensities, kind="stable")
    coord = np.transpose(coord)[idx_maxsort]
    max_out = int(num_peaks) if np.isfinite(num_peaks) else None
    coord = ensure_spacing_literal(coord, spacing=min_distance, p_norm=np.inf, max_out=max_out)
    if len(coord) > num_peaks:
        coord = coord[:num_peaks]
    return coord.astype(np.int64).reshape(-1, 2)


def polygon_cv2(r, c, shape=None):
    """Lattice points of skimage.draw.polygon(r, c, shape) for integer vertices, decided by OpenCV's exact integer
    point-in-polygon test (>= 0: inside or on the boundary).  Same row-major order as skimage's output."""
    import cv2

    r = np.asarray(r, np.int64); c = np.asarray(c, np.int64)
    minr, maxr = max(0, int(r.min())), int(r.max())
    minc, maxc = max(0, int(c.min())), int(c.max())
    if shape is not None:
        maxr = min(shape[0] - 1, maxr)
        maxc = min(shape[1] - 1, maxc)
    contour = np.stack([c, r], 1).astype(np.int32).reshape(-1, 1, 2)  # OpenCV points are (x=col, y=row)
    rr, cc = [], []
    for ri in range(minr, maxr + 1):
        for ci in range(minc, maxc + 1):
            if cv2.pointPolygonTest(contour, (float(ci), float(ri)), False) >= 0:
                rr.append(ri); cc.append(ci)
    return np.asarray(rr, np.int64), np.asarray(cc, np.int64)
