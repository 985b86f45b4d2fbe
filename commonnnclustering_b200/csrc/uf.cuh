// uf.cuh -- lock-free union-find on an int32 parent array (shared by cn.cu and packed.cu).
// parent[x] = x for roots, -1 for points that are not core; hooking always puts the larger root under the smaller
// one, so the final root of a component is its smallest member whatever the schedule.
#pragma once

__device__ __forceinline__ int ld_parent(const int *parent, int x) { return __ldcg(parent + x); }

// find with path halving; parents only ever move towards smaller indices
__device__ __forceinline__ int uf_find(int *parent, int x) {
    int p = ld_parent(parent, x);
    while (p != x) {
        int gp = ld_parent(parent, p);
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ int uf_find_ro(const int *parent, int x) {
    int p = ld_parent(parent, x);
    while (p != x) {
        x = p;
        p = ld_parent(parent, x);
    }
    return x;
}

// returns true when the call linked two previously separate trees
__device__ __forceinline__ bool uf_union(int *parent, int a, int b) {
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    while (ra != rb) {
        if (ra < rb) { int t = ra; ra = rb; rb = t; }
        int old = atomicCAS(parent + ra, ra, rb);
        if (old == ra) return true;
        ra = uf_find(parent, ra);
        rb = uf_find(parent, rb);
    }
    return false;
}

