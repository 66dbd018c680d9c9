// smooth3d16.cuh - block-Jacobi smoother specialised for the flagship geometry D = 3, N = 16
// (one patch = 256 pencils = one 256-thread CTA).  Included by kernels.cuh.
//
// Same arithmetic as the generic smooth_kernel (SchurHelper::solveWithSolution, SchurHelper.h:319-331:
// interface values + StarPatchOp::addInterfaceToRHS, StarPatchOp.h:185-203 + the DST patch solve of
// PatchSolvers/FftwPatchSolver.h:174-206 / DftPatchSolver.h:173-216); what changes is how the tile moves:
//   * tile layout x + 18 y + 290 z (rows 16-byte aligned): z- and y-pencil accesses are conflict-free
//     64-bit, x-row accesses are conflict-free 128-bit (half the LDS/STS count), column accesses
//     (x fixed, y varying) are 2-way conflicted at worst;
//   * transform order z, x, (y forward, eigenvalues, y inverse), x, z with warp w owning the rows
//     y in {2w, 2w+1}: the z<->x transposes stay inside a warp (__syncwarp instead of a CTA barrier),
//     leaving two CTA barriers around the y phase (+1 when gamma has to be subtracted first);
//   * f streams in with 16-byte cp.async into the second buffer, completion tracked by an mbarrier
//     (cp.async.mbarrier.arrive), so "the tile has landed" costs no CTA barrier;
//   * WRITE_U = false (sweeps whose u is only ever seen through its boundary slices: every sweep but
//     the last of a level visit in the fused cycle): the last inverse transform is evaluated in full
//     only for the 60 pencils on the patch boundary (warps 0-1); the other pencils compute just their
//     two z-face values (16 DFMA instead of 116) and nothing but the face buffer is written.
#pragma once

namespace tgpu
{
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void     cp_async16(double *smem_dst, const double *gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// the executing thread's arrival fires once all of its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar)
{
	asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
	"{\n"
	".reg .pred p;\n"
	"TGPU_MBAR_WAIT:\n"
	"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	"@p bra TGPU_MBAR_DONE;\n"
	"bra TGPU_MBAR_WAIT;\n"
	"TGPU_MBAR_DONE:\n"
	"}\n" ::"r"(smem_u32(bar)),
	"r"(parity)
	: "memory");
}

constexpr int    S16_ROW = 18, S16_PL = 290, S16_TILE = 16 * S16_PL;
constexpr size_t smooth3d16_smem_bytes() { return sizeof(double) * 2 * S16_TILE; }

template <bool ZERO_GUESS, bool EMIT, bool PROLONG, bool WRITE_U>
__global__ void __launch_bounds__(TGPU_THREADS, 3)
smooth3d16_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ f, double *__restrict__ u,
                  const double *__restrict__ Fin, double *__restrict__ Fout, const double *__restrict__ eig,
                  const double *__restrict__ uc)
{
	constexpr int N = 16, ROW = S16_ROW, PL = S16_PL;
	using G = Geo<3, 16>;
	static_assert(WRITE_U || EMIT, "a sweep must produce something");
	extern __shared__ __align__(16) double smem[];
	__shared__ uint64_t                    mbar[2];
	const int t = threadIdx.x, lo = t & 15, hi = t >> 4;
	const int npatch = P - p0;
	if (t == 0) {
		mbar_init(&mbar[0], TGPU_THREADS);
		mbar_init(&mbar[1], TGPU_THREADS);
	}
	__syncthreads();
	Mags<N> mg;
	mg.load();

	auto prefetch = [&](int g, int b) {
		const double *src = f + (size_t) (p0 + g) * G::NC;
		double *      dst = smem + b * S16_TILE;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const int c = t + TGPU_THREADS * i, row = c >> 3; // 16-byte chunk c of the patch, row = y + 16 z
			cp_async16(dst + (row & 15) * ROW + (row >> 4) * PL + (c & 7) * 2, src + c * 2);
		}
		cp_async_mbar_arrive(&mbar[b]);
	};

	int g = blockIdx.x;
	if (g < npatch) prefetch(g, 0);
	for (int it = 0; g < npatch; g += gridDim.x, it++) {
		const int        b      = it & 1;
		double *         S      = smem + b * S16_TILE;
		const int        p      = p0 + g;
		const int        gn     = g + gridDim.x;
		const PatchMeta &pm     = meta[p];
		const double     h2     = pm.h2;
		double           gam[6] = {0, 0, 0, 0, 0, 0}; // (2/h^2) gamma of entry t on each side (0: no neighbour)
		if (!ZERO_GUESS) {
			const double cfac = 2.0 * pm.inv_h2;
			int          ty[G::S];
			double       own[G::S], gm[G::S];
			const FaceVals<3, N, PROLONG ? FV_PROLONG : FV_PLAIN> fvals{Fin, uc, meta};
			gamma_all_sides(pm, p, t, fvals, ty, own, gm);
#pragma unroll
			for (int s = 0; s < G::S; s++) gam[s] = (ty[s] == NBR_NONE) ? 0.0 : cfac * gm[s];
		}
		mbar_wait(&mbar[b], (it >> 1) & 1); // every thread's cp.async of this tile has landed
		if (!ZERO_GUESS) {
			// x faces: entry t = (y, z) = (lo, hi).  The threads that touch the same edge cell through a
			// y face, entry (x, z), share z and hence the warp: a warp-level sync orders the two updates.
			double *r = S + lo * ROW + hi * PL;
			r[0] -= gam[0];
			r[N - 1] -= gam[1];
			__syncwarp();
			double *c = S + lo + hi * PL;
			c[0] -= gam[2];
			c[(N - 1) * ROW] -= gam[3];
			__syncthreads();
		}
		double v[N];
		{ // z forward: pencil (x, y) = (lo, hi)
			double *q = S + lo + hi * ROW;
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = q[k * PL];
			if (!ZERO_GUESS) {
				v[0] -= gam[4];
				v[N - 1] -= gam[5];
			}
			dst2_forward<N>(v, mg);
#pragma unroll
			for (int k = 0; k < N; k++) q[k * PL] = v[k];
		}
		__syncwarp(); // rows (y, k_z) with y in {2w, 2w+1} were produced by this warp
		double2 *rowp = reinterpret_cast<double2 *>(S + hi * ROW + lo * PL); // row (y, k_z) = (hi, lo)
		{
#pragma unroll
			for (int j = 0; j < N / 2; j++) {
				const double2 d = rowp[j];
				v[2 * j]        = d.x;
				v[2 * j + 1]    = d.y;
			}
			dst2_forward<N>(v, mg);
#pragma unroll
			for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
		}
		__syncthreads();
		// every thread is past the previous iteration: the other buffer may be refilled
		if (gn < npatch) {
			prefetch(gn, b ^ 1);
			if (!ZERO_GUESS) prefetch_faces_l2<3, N>(meta, p0 + gn, t, Fin);
		}
		{ // y forward, eigenvalues, y inverse: pencil (k_x, k_z) = (lo, hi)
			double *q = S + lo + hi * PL;
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = q[k * ROW];
			dst2_forward<N>(v, mg);
			const double *er = eig + t; // eig[k_y * 256 + k_x + 16 k_z] (the table is symmetric in the axes)
#pragma unroll
			for (int k = 0; k < N; k++) v[k] *= h2 * __ldg(er + k * G::M);
			dst3_inverse<N>(v, mg);
#pragma unroll
			for (int k = 0; k < N; k++) q[k * ROW] = v[k];
		}
		__syncthreads();
		{ // x inverse
#pragma unroll
			for (int j = 0; j < N / 2; j++) {
				const double2 d = rowp[j];
				v[2 * j]        = d.x;
				v[2 * j + 1]    = d.y;
			}
			dst3_inverse<N>(v, mg);
#pragma unroll
			for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
		}
		if (WRITE_U) {
			__syncwarp();
			double *q = S + lo + hi * ROW; // z inverse: pencil (x, y) = (lo, hi)
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = q[k * PL];
			dst3_inverse<N>(v, mg);
			double *up = u + (size_t) p * G::NC + t;
#pragma unroll
			for (int k = 0; k < N; k++) up[k * G::M] = v[k];
			if (EMIT) {
				double *Fp       = Fout + (size_t) p * G::S * G::M;
				Fp[4 * G::M + t] = v[0];
				Fp[5 * G::M + t] = v[N - 1];
				if (lo == 0) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[0 * G::M + k * N + hi] = v[k];
				}
				if (lo == N - 1) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[1 * G::M + k * N + hi] = v[k];
				}
				if (hi == 0) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[2 * G::M + k * N + lo] = v[k];
				}
				if (hi == N - 1) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[3 * G::M + k * N + lo] = v[k];
				}
			}
		} else {
			__syncthreads();
			// Only the boundary-cell slices of u are needed.  Warp 0: the x = 0 and x = 15 columns, warp 1:
			// the y = 0 and y = 15 rows (+ 4 interior pencils): full inverse transform; warps 2-7: the other
			// 192 interior pencils, z-face values only:
			//   u_0 = sum_j T[0][j] v_j,  u_15 = sum_j (-1)^j T[0][j] v_j,
			//   T[0][j] = sin(pi (j+1) / 32), T[0][15] = 1/2   (DftPatchSolver.h:269-281)
			double *  Fp   = Fout + (size_t) p * G::S * G::M;
			const int lane = t & 31;
			if (t < 64) {
				int x, y;
				if (t < 32) {
					x = (lane < 16) ? 0 : N - 1;
					y = lane & 15;
				} else if (lane < 28) {
					x = 1 + lane % 14;
					y = (lane < 14) ? 0 : N - 1;
				} else {
					x = 11 + (lane - 28); // interior pencils 192..195 = (11..14, 14)
					y = 14;
				}
				const double *q = S + x + y * ROW;
#pragma unroll
				for (int k = 0; k < N; k++) v[k] = q[k * PL];
				dst3_inverse<N>(v, mg);
				Fp[4 * G::M + x + N * y] = v[0];
				Fp[5 * G::M + x + N * y] = v[N - 1];
				if (x == 0) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[0 * G::M + k * N + y] = v[k];
				}
				if (x == N - 1) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[1 * G::M + k * N + y] = v[k];
				}
				if (y == 0) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[2 * G::M + k * N + x] = v[k];
				}
				if (y == N - 1) {
#pragma unroll
					for (int k = 0; k < N; k++) Fp[3 * G::M + k * N + x] = v[k];
				}
			} else {
				const int     idx = t - 64, x = 1 + idx % 14, y = 1 + idx / 14;
				const double *q   = S + x + y * ROW;
				double        E = 0.0, O = 0.0;
#pragma unroll
				for (int j = 0; j < N; j += 2) {
					E = fma(mg.sinq(j + 1), q[j * PL], E);
					O = fma((j + 1 == N - 1) ? 0.5 : mg.sinq(j + 2), q[(j + 1) * PL], O);
				}
				Fp[4 * G::M + x + N * y] = E + O;
				Fp[5 * G::M + x + N * y] = E - O;
			}
		}
	}
}
} // namespace tgpu
