"""CPU oracle for the ingest indexing (od_pipe.py:214-247, 253-288) -- NumPy restatement.

TEST INFRASTRUCTURE ONLY (same rules as ba_oracle.py).  PARITY PINNED: tests/test_indexing.py compares these functions
with `tests/golden/seq_a.npz` / `seq_b.npz`, which hold the outputs of the reference's own `read_detections` and
`remove_elems` (tests/golden/make_golden_streaming.py).  The product computes the same integers with device kernels
(vinsat_b200/csrc/ingest.cu).
"""
import numpy as np


def index_detections(frames, n_orbit):
    """od_pipe.py:214-228,242-247: unique frames, knot frames at multiples of 1000 s, slot of every detection.
    frames = detections[:, 0] (sorted).  Returns (time_idx with knots, ii)."""
    frames = np.asarray(frames)
    uniq, counts = np.unique(frames, return_counts=True)
    time_idx = uniq.astype(np.int64)
    filler_idx = time_idx.min() // 1000 + 1
    filler_offset = 0
    time_idx_new, slots = [], []
    for i, tidx in enumerate(time_idx):                     # :218-228
        if tidx == filler_idx * 1000:
            filler_idx += 1
        while tidx > filler_idx * 1000:
            time_idx_new.append(filler_idx * 1000)
            filler_idx += 1
            filler_offset += 1
        time_idx_new.append(tidx)
        slots.append(i + filler_offset)
    ii = np.repeat(np.array(slots, dtype=np.int64), counts)
    if time_idx[-1] < n_orbit:                              # :242-245
        while filler_idx * 1000 < (n_orbit // 1000) * 1000 + 1:
            time_idx_new.append(filler_idx * 1000)
            filler_idx += 1
    return np.array(time_idx_new, dtype=np.int64), ii


def remove_elems_index(mask, ii, time_idx):
    """od_pipe.py:253-288 as an exclusive prefix sum over the kept-frame mask: frames kept = frames with a
    surviving observation or knots; ii_new[k] = ii_old[k] - #{dropped frames < ii_old[k]}.
    Returns (ii_new, time_idx_new, keep)."""
    mask = np.asarray(mask, dtype=bool)
    ii = np.asarray(ii, dtype=np.int64)
    time_idx = np.asarray(time_idx, dtype=np.int64)
    ii_old = ii[mask]
    keep = np.zeros(time_idx.shape[0], dtype=bool)
    keep[np.unique(ii_old)] = True
    keep |= (time_idx % 1000 == 0)
    dropped = ~keep
    if len(ii_old):
        dropped[int(ii_old.max()) + 1:] = False           # the reference only walks i <= ii_old.max() (:272)
    else:
        dropped[:] = False
    shift = np.concatenate([[0], np.cumsum(dropped)[:-1]]).astype(np.int64)
    return ii_old - shift[ii_old], time_idx[keep], keep
