"""Host-side model of how the device states the reference's breadth-first search (DESIGN.md 1 / SURVEY Appendix A): the
clusters of `FitterExtBFS._fit` (_fit.pyx:408-491) are the connected components of the graph whose vertices are the
points with enough neighbours and whose edges are the pairs (i, j in N(i)) that pass the common-neighbour check; the BFS
numbers them in the order of their smallest member.  The CUDA path builds the same components with a union-find over the
edges in WHATEVER order the warps reach them, in rounds (adjacent entry, far entries, everything else), skipping edges
whose ends already share a root -- this file checks on the CPU, against the oracle's BFS, that neither the order nor the
rounds nor the skipping can change a label, and that the one-key ranking `(N - size : smallest member)` the device sorts
by reproduces `Labels.sort_by_size` (_types.pyx:212-286) with its `member_cutoff` / `max_clusters` rules.  No GPU."""
import numpy as np
import pytest

from oracle import port


def find(parent, a):
    while parent[a] != a:
        parent[a] = parent[parent[a]]
        a = parent[a]
    return a


def components_by_union_find(row_ptr, col_idx, n_member_cutoff, similarity_cutoff, rng, current_start=1):
    """labels before the size sort, the way cn.cu arrives at them; returns (labels, checks executed, candidate edges)."""
    n = row_ptr.size - 1
    deg = np.diff(row_ptr)
    core = deg > n_member_cutoff                     # `enough` is strict (_types.pyx:1408-1411)
    parent = np.arange(n)
    rows = [col_idx[row_ptr[i]:row_ptr[i + 1]] for i in range(n)]
    edges = [(i, int(j)) for i in range(n) if core[i] for j in rows[i] if j > i and core[j]]
    # rounds: A = the first entry above the diagonal of every row, B = the last entry of every 8th row, C = the rest,
    # each in a random order (the schedule of the warps)
    first_above = {}
    last = {}
    for e in edges:
        first_above.setdefault(e[0], e)
        last[e[0]] = e
    round_a = list(first_above.values())
    round_b = [e for i, e in last.items() if i % 8 == 0]
    checks, rejected = 0, set()
    for batch in (round_a, round_b, edges):
        for k in rng.permutation(len(batch)):
            i, j = batch[k]
            ri, rj = find(parent, i), find(parent, j)
            if ri == rj or (i, j) in rejected:
                continue                            # settled for free: the answer cannot change the components
            checks += 1
            rejected.add((i, j))                    # (an accepted edge is settled by its union from here on)
            if port.count_common(rows[i], rows[j]) >= similarity_cutoff:
                parent[max(ri, rj)] = min(ri, rj)
    root = np.array([find(parent, i) for i in range(n)])
    # numbering: components in the order of their smallest member (= the BFS's `init` loop); a point without enough
    # neighbours is noise, a core point nobody joined is a cluster of one
    labels = np.zeros(n, dtype=np.int64)
    smallest = {}
    for i in range(n):
        if core[i]:
            smallest.setdefault(root[i], i)
    order = {r: current_start + k for k, r in enumerate(sorted(smallest, key=smallest.get))}
    for i in range(n):
        if core[i]:
            labels[i] = order[root[i]]
    return labels, checks, len(edges)


def rank_by_one_key(labels, member_cutoff=None, max_clusters=None):
    """fused_rank_kernel's ordering: key (N - size : smallest member) ascending, survivors of member_cutoff only."""
    n = labels.size
    member_cutoff = 2 if member_cutoff is None else member_cutoff
    ids, first, size = np.unique(labels[labels > 0], return_index=True, return_counts=True)
    first = np.array([np.flatnonzero(labels == c)[0] for c in ids], dtype=np.int64)
    keep = size >= member_cutoff
    key = (n - size[keep]).astype(np.int64) * (n + 1) + first[keep]
    ranked = ids[keep][np.argsort(key, kind="stable")]
    if max_clusters is not None:
        ranked = ranked[:max_clusters]
    new = np.zeros(int(labels.max()) + 1, dtype=np.int64)
    new[ranked] = np.arange(1, ranked.size + 1)
    return new[labels]


def blobs(rng, n, d, k, spread):
    centres = rng.random((k, d)) * 10.0
    return centres[rng.integers(0, k, n)] + rng.standard_normal((n, d)) * spread


CASES = {
    "blobs_2d": (lambda rng: blobs(rng, 700, 2, 6, 0.35), 0.3, 4),
    "blobs_3d_sparse": (lambda rng: blobs(rng, 700, 3, 5, 0.6), 0.45, 2),
    "uniform_c0": (lambda rng: rng.random((500, 2)) * 6.0, 0.3, 0),
    "lattice_ties": (lambda rng: rng.integers(0, 30, (600, 2)).astype(np.float64) * 0.25, 0.5, 4),
    "high_cutoff_mostly_noise": (lambda rng: blobs(rng, 600, 2, 8, 0.5), 0.25, 12),
}


@pytest.mark.parametrize("name", list(CASES))
def test_union_find_in_any_order_equals_the_bfs(name):
    make, r, c = CASES[name]
    x = np.ascontiguousarray(make(np.random.default_rng(5)))
    p = port.make_parameters(r, c)
    row_ptr, col_idx = port.neighbours(x, p["radius_cutoff"], method="brute")
    want = port.fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], 1)
    assert want.max() >= 2 and (want == 0).any()          # clusters and noise both occur
    for seed in range(3):
        got, checks, n_edges = components_by_union_find(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"],
                                                        np.random.default_rng(seed))
        np.testing.assert_array_equal(got, want, err_msg=f"schedule {seed}")
        assert checks <= n_edges
    # the size sort as one ranking, with the reference's cut-offs
    for member_cutoff, max_clusters in [(None, None), (1, None), (5, None), (None, 2), (10, 3)]:
        ref = port.sort_by_size(want.copy(), member_cutoff, max_clusters)
        np.testing.assert_array_equal(rank_by_one_key(want, member_cutoff, max_clusters), ref,
                                      err_msg=f"member_cutoff={member_cutoff} max_clusters={max_clusters}")


def test_skipping_settled_edges_saves_most_checks():
    """The BFS tests every (point, unlabelled neighbour) pair it meets; the union-find needs one accepted edge per merge."""
    make, r, c = CASES["blobs_2d"]
    x = np.ascontiguousarray(make(np.random.default_rng(5)))
    p = port.make_parameters(r, c)
    row_ptr, col_idx = port.neighbours(x, p["radius_cutoff"], method="brute")
    _, stats = port.fit_bfs(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"], 1, return_stats=True)
    _, checks, n_edges = components_by_union_find(row_ptr, col_idx, p["n_member_cutoff"], p["similarity_cutoff"],
                                                  np.random.default_rng(0))
    assert checks < n_edges // 2 and checks < int(stats[1])
