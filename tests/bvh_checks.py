"""Structure checks on a downloaded LBVH, shared by the emulation and GPU tests."""
import numpy as np


def decode_child(code):
    code = int(code)
    if code >= 0:
        return ("node", code, 0)
    u = (~code) & 0xFFFFFFFF
    return ("leaf", u >> 3, (u & 7) + 1)


def check_bvh(nodes, order, keys, tri_v, leaf_size):
    """nodes: (n,16) float32 as stored; order: sorted position -> original triangle; keys sorted."""
    nb = len(order)
    assert np.all(keys[:-1] <= keys[1:]), "Morton keys not sorted"
    assert len(np.unique(order)) == nb, "triangle appears twice in the order"
    if nb < 2:
        return {"depth": 1, "leaves": nb}
    codes = nodes[:, 12:16].copy().view(np.int32)
    tri = tri_v.reshape(-1, 3, 3)[order]
    tlo, thi = tri.min(axis=1), tri.max(axis=1)
    seen = np.zeros(nb, np.int32)
    max_depth = 0
    n_leaves = 0
    # iterative DFS from the root carrying the box the parent stored for this child
    stack = [(0, None, None, 1)]
    visited_nodes = set()
    while stack:
        node, plo, phi, depth = stack.pop()
        assert node not in visited_nodes, "node reachable twice"
        visited_nodes.add(node)
        max_depth = max(max_depth, depth)
        n = nodes[node]
        first, last = int(codes[node, 2]), int(codes[node, 3])
        boxes = [(np.array([n[0], n[2], n[8]]), np.array([n[1], n[3], n[9]])),
                 (np.array([n[4], n[6], n[10]]), np.array([n[5], n[7], n[11]]))]
        covered = 0
        for ci in range(2):
            kind, a, cnt = decode_child(codes[node, ci])
            lo, hi = boxes[ci]
            if plo is not None:
                assert np.all(lo >= plo - 1e-3 * (1 + np.abs(plo))) and np.all(hi <= phi + 1e-3 * (1 + np.abs(phi))), \
                    "child box not inside the parent's box"
            if kind == "leaf":
                assert 1 <= cnt <= max(leaf_size, 1)
                assert first <= a and a + cnt - 1 <= last
                seen[a:a + cnt] += 1
                n_leaves += 1
                covered += cnt
                assert np.all(tlo[a:a + cnt] >= lo) and np.all(thi[a:a + cnt] <= hi), "leaf box does not contain its triangles"
            else:
                cf, cl = int(codes[a, 2]), int(codes[a, 3])
                assert first <= cf and cl <= last
                covered += cl - cf + 1
                stack.append((a, lo, hi, depth + 1))
        assert covered == last - first + 1, "children do not partition the node's range"
    assert np.all(seen == 1), "some triangle is not reachable exactly once"
    return {"depth": max_depth, "leaves": n_leaves, "nodes_reachable": len(visited_nodes)}
