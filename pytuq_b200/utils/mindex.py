"""Multi-index tables (host side; they define the column order of the design matrix).

Counterpart of the reference's ``utils/mindex.py``.  The tables are tiny compared with
the sample axis, stay on the host and are uploaded once per plan.  Row orders are
identical to the reference's (they fix the meaning of every coefficient vector):
``get_mi`` follows utils/mindex.py:29-62, the front-growing routines utils/mindex.py:137-242.
"""
import math

import numpy as np


def get_npc(ord, dim):
    """Number of terms of total order <= ord in dim variables, (ord+dim)!/(ord! dim!)
    (reference: utils/mindex.py:7-26)."""
    return math.comb(int(ord) + int(dim), int(ord))


def get_mi(ord, dim):
    """Total-order multi-index array of shape (get_npc(ord, dim), dim), dtype int.

    Graded by total order.  Inside the block of order q the rows are produced dimension by
    dimension: rows that add one to dimension j are built from the trailing ``tail[j]`` rows
    of the order q-1 block, where ``tail[j]`` counts the order q-1 rows whose first non-zero
    dimension is >= j.  This reproduces the reference's row order (utils/mindex.py:29-62).
    """
    assert dim > 0
    blocks = [np.zeros((1, dim), dtype=int)]
    if ord >= 1:
        blocks.append(np.eye(dim, dtype=int))
        tail = np.ones(dim, dtype=int)
        for _ in range(2, ord + 1):
            prev = blocks[-1]
            tail = np.cumsum(tail[::-1])[::-1].copy()
            pieces = []
            for j in range(dim):
                piece = prev[prev.shape[0] - tail[j]:].copy()
                piece[:, j] += 1
                pieces.append(piece)
            blocks.append(np.concatenate(pieces, axis=0))
    return np.concatenate(blocks, axis=0)


def encode_mindex(mindex):
    """Sparse (dimension list, order list) form of each row, dimensions 1-based
    (reference: utils/mindex.py:65-96)."""
    npc = mindex.shape[0]
    print(f"Multiindex has {npc} terms")
    pairs = []
    for row in mindex:
        nz = np.flatnonzero(row)
        if nz.size == 0:
            pairs.append(([0], [0]))
        else:
            pairs.append(((nz + 1).tolist(), row[nz].tolist()))
    return pairs


def micf_join(mindex_list, cfs_list):
    """Union of several multi-index sets with the coefficients re-indexed onto it
    (reference: utils/mindex.py:99-133).  Returns (mindex_all (K,s), cfs_all (nout,K));
    the union is in np.unique (lexicographic) order and always contains the zero row."""
    nout = len(mindex_list)
    assert nout == len(cfs_list)
    sdim = mindex_list[0].shape[1]
    stacked = np.concatenate(list(mindex_list) + [np.zeros((1, sdim), dtype=int)])
    mindex_all = np.unique(stacked, axis=0)
    lookup = {tuple(row): k for k, row in enumerate(mindex_all.tolist())}
    cfs_all = np.zeros((nout, mindex_all.shape[0]))
    for j, (mi, cf) in enumerate(zip(mindex_list, cfs_list)):
        seen = set()
        for r, row in enumerate(mi.tolist()):
            k = lookup[tuple(row)]
            if k not in seen:           # first occurrence wins, as in the reference
                seen.add(k)
                cfs_all[j, k] = cf[r]
    return mindex_all, cfs_all


def _grow_front(mindex, need_all_parents):
    npc, ndim = mindex.shape
    have = set(map(tuple, mindex.tolist()))
    added, added_set, front = [], set(), []
    for row in mindex.tolist():
        in_front = False
        for j in range(ndim):
            cand = list(row)
            cand[j] += 1
            tc = tuple(cand)
            if tc in have:
                continue
            if need_all_parents:
                ok = True
                for k in range(ndim):
                    if cand[k] != 0:
                        par = list(cand)
                        par[k] -= 1
                        if tuple(par) not in have:
                            ok = False
                            break
                if not ok:
                    continue
            if tc not in added_set:
                added_set.add(tc)
                added.append(cand)
            if not in_front:
                front.append(row)
                in_front = True
    mindex_add = np.array(added, dtype=int).reshape(-1, ndim)
    mindex_f = np.array(front, dtype=int).reshape(-1, ndim)
    mindex_new = np.vstack((mindex, mindex_add))
    return [mindex_new, mindex_add, mindex_f]


def mi_addfront_cons(mindex):
    """Adds the admissible forward neighbours (all backward neighbours present)
    (reference: utils/mindex.py:137-196).  Returns [new, added, front]."""
    return _grow_front(mindex, True)


def mi_addfront(mindex):
    """Adds every forward neighbour not yet in the set (reference: utils/mindex.py:202-242).
    Returns [new, added, front]."""
    return _grow_front(mindex, False)
