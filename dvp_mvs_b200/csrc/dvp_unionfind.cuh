// dvp_unionfind.cuh — lock-free union-find over a pixel grid, shared by the visibility restoration (row N1,
// dvp_kernels_post.cu) and the Canny hysteresis of the depth-edge prior (row N4, dvp_kernels_edge.cu).
#pragma once
#include <cuda_runtime.h>

namespace dvp {

// ---- union-find on parent[] (parent[p] == p: root; -1: pixel not in any region) -----------------------
// Links only ever change at roots and only downwards (atomicMin), so a stale read is still an ancestor and
// the atomic's return value decides; reads go through volatile to pick up other SMs' links early.
__device__ __forceinline__ int uf_find(const int* parent, int a) {
	const volatile int* p = parent;
	int up = p[a];
	while (up != a) { a = up; up = p[a]; }
	return a;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
	for (;;) {
		a = uf_find(parent, a);
		b = uf_find(parent, b);
		if (a == b) return;
		if (a > b) { const int t = a; a = b; b = t; }   // a < b: hang b under a
		const int old = atomicMin(&parent[b], a);
		if (old == b) return;                            // b was still a root: linked
		b = old;                                          // somebody else re-parented b first: continue from there
	}
}

}  // namespace dvp
