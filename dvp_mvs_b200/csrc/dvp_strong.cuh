// dvp_strong.cuh — device code of the strong-pixel path: random plane generation, view selection,
// red-black propagation and refinement (reference APD.cu:501-669, 1115-1194, 1311-1383, 2010-2141, 2462-2567,
// 2725-2736).  One thread owns one pixel; everything that the reference keeps in per-thread local-memory
// arrays (cost_array[8][32], positions[8], sampling_probs[32] ...) lives in shared memory, column-per-thread,
// so the only local memory left is the 20-entry view-direction table of the normal sampler.
#pragma once
#include "dvp_ncc.cuh"
#include "dvp_launch.h"

namespace dvp {

// Packed 4-bit view-weight counters (each in 0..15, 15 draws), 32 views in 4 words.
struct ViewWeights {
	uint32_t w[4];
	__device__ __forceinline__ void clear() { w[0] = w[1] = w[2] = w[3] = 0; }
	__device__ __forceinline__ int get(int v) const { return (w[v >> 3] >> ((v & 7) * 4)) & 15; }
	__device__ __forceinline__ void inc(int v) { w[v >> 3] += 1u << ((v & 7) * 4); }
	__device__ __forceinline__ void store(uint8_t* dst) const {  // dst: 32 bytes, 16 B aligned
		uint32_t o[8];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const uint32_t x = w[q];
			o[2 * q] = (x & 15) | (((x >> 4) & 15) << 8) | (((x >> 8) & 15) << 16) | (((x >> 12) & 15) << 24);
			o[2 * q + 1] = ((x >> 16) & 15) | (((x >> 20) & 15) << 8) | (((x >> 24) & 15) << 16) | (((x >> 28) & 15) << 24);
		}
		reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
		reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
	}
	__device__ __forceinline__ void load(const uint8_t* src) {
		const uint4 a = reinterpret_cast<const uint4*>(src)[0];
		const uint4 b = reinterpret_cast<const uint4*>(src)[1];
		const uint32_t o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const uint32_t lo = o[2 * q], hi = o[2 * q + 1];
			w[q] = (lo & 15) | (((lo >> 8) & 15) << 4) | (((lo >> 16) & 15) << 8) | (((lo >> 24) & 15) << 12) |
			       ((hi & 15) << 16) | (((hi >> 8) & 15) << 20) | (((hi >> 16) & 15) << 24) | (((hi >> 24) & 15) << 28);
		}
	}
};

// reference GenerateRandomNormal_YZL, APD.cu:501-588 (bugs B4, B12 reproduced).
static __device__ __noinline__ float4 random_normal(const KArgs& a, int px, int py, Rng& rng, float depth, uint32_t sel) {
	float3 vdir[20];
	{
		const float4 v0 = get_view_direction(a.ref, px, py, depth);
		vdir[0] = make_float3(v0.x, v0.y, v0.z);
	}
	int index = 1;
	for (int v = 0; v < a.S; ++v) {
		if (!is_set(sel, v)) continue;
		if (index >= 20) break;  // the reference overruns view_direction[20] here (B12); we stop at its capacity
		const ViewConst& vc = a.views[v];
		const float3 fwd = point_to_world((float)px, (float)py, depth, a.ref.K, a.ref.R, a.ref.c);
		float2 src_pt; float src_d;
		project_on_camera(fwd, vc.sK, vc.sR, vc.st, src_pt, src_d);
		const int sx = (int)((int)src_pt.x + 0.5f), sy = (int)((int)src_pt.y + 0.5f);
		float src_depth = 1.0f;
		if (a.prm.geom_consistency) {
			// out-of-image projections leave src_depth uninitialised in the reference (UB); 1.0 gives the
			// same direction as any positive value because the direction is normalised.
			if (sx >= 0 && sx < a.W && sy >= 0 && sy < a.H)
				{ src_depth = tex2D<float>(a.tex_depth[v + 1], (int)src_pt.x + 0.5f, (int)src_pt.y + 0.5f); DVP_COUNT(a, 1); }
		}
		const dvp_camera& sc = a.cams[v + 1];
		const float4 d = get_view_direction(sc, sx, sy, src_depth);
		const float dir[3] = {d.x, d.y, d.x};  // B4: x used for the z component
		float Rf[3];
		Rf[0] = vc.R_c[0] * dir[0] + vc.R_c[1] * dir[1] + vc.R_c[2] * dir[2];
		Rf[1] = vc.R_c[3] * dir[0] + vc.R_c[4] * dir[1] + vc.R_c[5] * dir[2];
		Rf[2] = vc.R_c[6] * dir[0] + vc.R_c[7] * dir[1] + vc.R_c[7] * dir[2];  // B4: A[7] twice
		const float norm = sqrt(Rf[0] * Rf[0] + Rf[1] * Rf[1] + Rf[2] * Rf[2]);
		vdir[index++] = make_float3(Rf[0] / norm, Rf[1] / norm, Rf[2] / norm);
	}
	int times = 200;
	float4 normal;
	// up to 5 directions (the reference view + 4 selected sources, every schedule of the reference) are tested from registers:
	// the rejection loop below runs up to 200 times and would otherwise re-read them from the stack each time.  The test has
	// no side effects, so evaluating all of them instead of stopping at the first failure gives the same answer.
	constexpr int kRegDirs = 5;
	float3 rd[kRegDirs];
#pragma unroll
	for (int i = 0; i < kRegDirs; ++i) rd[i] = i < index ? vdir[i] : make_float3(0.f, 0.f, 0.f);
	while (times > 0) {
		float q1 = 1.0f, q2 = 1.0f, s = 2.0f;
		while (s >= 1.0f) {
			q1 = 2.0f * rng.uniform() - 1.0f;
			q2 = 2.0f * rng.uniform() - 1.0f;
			s = q1 * q1 + q2 * q2;
		}
		const float sq = sqrt(1.0f - s);
		normal.x = 2.0f * q1 * sq;
		normal.y = 2.0f * q2 * sq;
		normal.z = 1.0f - 2.0f * s;
		normal.w = 0;
		bool satisfy = true;
#pragma unroll
		for (int i = 0; i < kRegDirs; ++i) {
			const float d = normal.x * rd[i].x + normal.y * rd[i].y + normal.z * rd[i].z;
			if (i < index && d > 0.0f) satisfy = false;
		}
		for (int i = kRegDirs; i < index && satisfy; i++) {
			const float d = normal.x * vdir[i].x + normal.y * vdir[i].y + normal.z * vdir[i].z;
			if (d > 0.0f) { satisfy = false; break; }
		}
		if (satisfy) break;
		times--;
	}
	normalize3(&normal);
	return normal;
}

// reference GeneratePerturbedNormal, APD.cu:617-661.  B3: the perturbed normal is discarded — the function
// returns normalize(normal) — but the RNG draws and the loop exit test must still be reproduced.
static __device__ __noinline__ float4 perturbed_normal(const KArgs& a, int px, int py, const float4 normal, Rng& rng, const float perturbation) {
	const float4 view = get_view_direction(a.ref, px, py, 1.0f);
	float4 out = normal;
	int times = 200;
	while (times > 0) {
		const float a1 = (rng.uniform() - 0.5f) * perturbation;
		const float a2 = (rng.uniform() - 0.5f) * perturbation;
		const float a3 = (rng.uniform() - 0.5f) * perturbation;
		const float sin_a1 = sin(a1), sin_a2 = sin(a2), sin_a3 = sin(a3);
		const float cos_a1 = cos(a1), cos_a2 = cos(a2), cos_a3 = cos(a3);
		float R[9];
		R[0] = cos_a2 * cos_a3;
		R[1] = cos_a3 * sin_a1 * sin_a2 - cos_a1 * sin_a3;
		R[2] = sin_a1 * sin_a3 + cos_a1 * cos_a3 * sin_a2;
		R[3] = cos_a2 * sin_a3;
		R[4] = cos_a1 * cos_a3 + sin_a1 * sin_a2 * sin_a3;
		R[5] = cos_a1 * sin_a2 * sin_a3 - cos_a3 * sin_a1;
		R[6] = -sin_a2;
		R[7] = cos_a2 * sin_a1;
		R[8] = cos_a1 * cos_a2;
		float4 np;
		np.x = R[0] * normal.x + R[1] * normal.y + R[2] * normal.z;
		np.y = R[3] * normal.x + R[4] * normal.y + R[5] * normal.z;
		np.z = R[6] * normal.x + R[7] * normal.y + R[8] * normal.z;
		if (np.x * view.x + np.y * view.y + np.z * view.z < 0.0f) break;
		times--;
	}
	normalize3(&out);
	return out;
}

// Weighted multi-view cost of one hypothesis over the views with a positive weight.
// (the reference evaluates all S views and multiplies the unused ones by 0 — identical sum.)
// `reject_at`: the caller only asks whether the result is < reject_at (refinement, APD.cu:1373).  Every term is
// >= 0 (an NCC lies in [0, 2]), float addition of a non-negative term never lowers the sum and the scaling by
// 1 / weight_norm is monotone, so as soon as the scaled partial sum is no longer below reject_at the full cost cannot
// be either: the remaining views are not evaluated.  The value returned then only has to fail the same test.
__device__ __forceinline__ float weighted_cost(const KArgs& a, int px, int py, const float4 pl, const RefPatch& rp,
                                               const float2* wt, int stride, const ViewWeights& vw, float weight_norm, float reject_at) {
	float acc = 0.0f;
	for (int v = 0; v < a.S; ++v) {
		const int wv = vw.get(v);
		if (wv > 0) {
			const float c = ncc_cost<kSweepRB, kSweepRW>(a, a.views[v], a.tex_img[v + 1], px, py, pl, rp, wt, stride);
			acc += wv * c;
#ifdef DVP_EARLY_REJECT   // measured on B200: 10 % SLOWER than evaluating every view (the data-dependent loop exit breaks the fetch batching); off
			if (!(acc / weight_norm < reject_at)) break;
#endif
		}
	}
	acc /= weight_norm;
	return acc;
}

// reference PlaneHypothesisRefinementStrong, APD.cu:1311-1383.
__device__ __forceinline__ void refine_strong(const KArgs& a, int px, int py, float4* plane, float* depth, float* cost, Rng& rng,
                                              const ViewWeights& vw, float weight_norm, uint32_t sel_now,
                                              const RefPatch& rp, const float2* wt, int stride) {
	const float depth_perturbation = 0.02f;
	const float normal_perturbation = 0.02f;
	const float depth_min = a.prm.depth_min, depth_max = a.prm.depth_max;

	const float depth_rand = rng.uniform() * (depth_max - depth_min) + depth_min;
	const float4 plane_rand = random_normal(a, px, py, rng, *depth, sel_now);
	float depth_perturbed = *depth;
	const float depth_min_perturbed = (1 - depth_perturbation) * depth_perturbed;
	const float depth_max_perturbed = (1 + depth_perturbation) * depth_perturbed;
	do {
		depth_perturbed = rng.uniform() * (depth_max_perturbed - depth_min_perturbed) + depth_min_perturbed;
	} while (depth_perturbed < depth_min && depth_perturbed > depth_max);  // B14: never loops

	const float perturbation = normal_perturbation * 3.14159265358979323846;
	const float4 plane_pert = perturbed_normal(a, px, py, *plane, rng, perturbation);
	(void)perturbed_normal(a, px, py, *plane, rng, perturbation);  // second call: same result (B3), RNG still advances

	// hypotheses (depth, normal): (rand,cur) (cur,rand) (rand,rand) (cur,pert1) (cur,pert2) (pert,cur).
	// pert2 == pert1 bit for bit and is scored against the already updated cost with a strict '<', so
	// hypothesis 4 can never be accepted; it is skipped.
	const float depth0 = *depth;
	const float4 plane0 = *plane;
#pragma unroll 1
	for (int i = 0; i < 6; ++i) {
		if (i == 4) continue;
		float d; float4 t;
		switch (i) {
		case 0: d = depth_rand; t = plane0; break;
		case 1: d = depth0; t = plane_rand; break;
		case 2: d = depth_rand; t = plane_rand; break;
		case 3: d = depth0; t = plane_pert; break;
		default: d = depth_perturbed; t = plane0; break;
		}
		t.w = get_distance2origin(a.ref, px, py, d, t);
		const float depth_before = depth_from_plane(a.ref, t, px, py);
		if (!(depth_before >= depth_min && depth_before <= depth_max)) continue;   // rejected whatever it costs
		const float temp_cost = weighted_cost(a, px, py, t, rp, wt, stride, vw, weight_norm, *cost);
		if (temp_cost < *cost) {
			*depth = depth_before;
			*plane = t;
			*cost = temp_cost;
		}
	}
}

}  // namespace dvp
