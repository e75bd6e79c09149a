"""Host logic of the task-stream kernel (csrc/ewb_stream.cuh, schedule built in csrc/ewb_api.cu), on CPU through the test hook
ewb_debug_stream_schedule: every element and every node is handled exactly once, and a gather task depends only on element tasks
with a LOWER ticket — the property that makes the persistent kernel deadlock-free for any grid size."""
import ctypes as C

import numpy as np
import pytest

from edelweissfe_b200 import _lib, box_mesh


def _schedule(conn, n_node, order, chunk, delay, ept, npt):
    lib = _lib.load()
    n_el, nn = conn.shape
    max_tasks = n_el + n_node + 16
    tasks = np.zeros(2 * max_tasks, dtype=np.int32)
    gnodes = np.zeros(n_node, dtype=np.int32)
    target = np.zeros(n_el + 1, dtype=np.int32)
    o = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    n = lib.ewb_debug_stream_schedule(nn, n_el, n_node, conn.ctypes.data_as(C.c_void_p), None if o is None else o.ctypes.data_as(C.c_void_p),
                                      chunk, delay, ept, npt, tasks.ctypes.data_as(C.c_void_p), max_tasks, gnodes.ctypes.data_as(C.c_void_p),
                                      target.ctypes.data_as(C.c_void_p), target.size)
    assert n > 0, lib.ewb_last_error()
    return tasks[: 2 * n].reshape(n, 2), gnodes, target[: (n_el + chunk - 1) // chunk]


@pytest.mark.parametrize("chunk,delay,ept,npt", [(1, 0, 1, 1), (7, 1, 3, 5), (16, 2, 2, 8), (256, 12, 2, 8), (5, 1000, 4, 32)])
@pytest.mark.parametrize("ordered", [False, True])
def test_schedule_invariants(chunk, delay, ept, npt, ordered):
    coords, conn = box_mesh(4, 3, 5, elType="C3D20")
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    n_el, n_node = conn.shape[0], coords.shape[0] + 2  # two nodes without elements at the end
    order = np.random.default_rng(chunk + delay).permutation(n_el).astype(np.int32) if ordered else None
    tasks, gnodes, target = _schedule(conn, n_node, order, chunk, delay, ept, npt)
    pos_of = np.empty(n_el, dtype=np.int64)
    pos_of[order if ordered else np.arange(n_el)] = np.arange(n_el)
    chunk_of = pos_of // chunk
    assert sorted(gnodes.tolist()) == list(range(n_node))  # a permutation of the nodes
    ready = np.zeros(n_node, dtype=np.int64)
    np.maximum.at(ready, conn.reshape(-1), np.repeat(chunk_of, conn.shape[1]))
    kind, cnt, key, first = tasks[:, 0] & 1, (tasks[:, 0] >> 1) & 127, tasks[:, 0] >> 8, tasks[:, 1]
    seen_el = np.zeros(n_el, dtype=int)
    seen_node = np.zeros(n_node, dtype=int)
    done_tasks = np.zeros(target.size, dtype=int)  # element tasks of every chunk issued so far (ticket order)
    for t in range(tasks.shape[0]):
        if kind[t] == 0:
            p = np.arange(first[t], first[t] + cnt[t])
            assert (p // chunk == key[t]).all() and 1 <= cnt[t] <= ept
            seen_el[order[p] if ordered else p] += 1
            done_tasks[key[t]] += 1
        else:
            nodes = gnodes[first[t] : first[t] + cnt[t]]
            assert 1 <= cnt[t] <= npt
            seen_node[nodes] += 1
            assert (ready[nodes] <= key[t]).all()  # the nodes' elements lie in chunks <= key ...
            assert (done_tasks[: key[t] + 1] == target[: key[t] + 1]).all()  # ... whose element tasks all hold lower tickets
    assert (seen_el == 1).all() and (seen_node == 1).all()
    assert done_tasks.tolist() == target.tolist()
