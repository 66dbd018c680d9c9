// patch3d32.cuh - kernels for D = 3, N = 32 patches (BASELINE config D).  Included by kernels.cuh.
//
// A 32^3 patch is 256 KB of fp64: more than one SM's shared memory, so the one-CTA-one-tile scheme of the
// smaller sizes does not apply.
//   smooth3d32c_kernel  the smoother the library uses: a thread-block cluster of two CTAs holds the whole patch in the
//                       shared memory of two SMs; x and y are transformed, z is a two-sided tridiagonal elimination split
//                       over the pair (see the comment above the kernel)
//   smooth3d32n_kernel  levels with Neumann domain sides: general dense-transform path through an L2-resident scratch block
//   smooth3d32_kernel   the first version (PR 1), kept as the cross-check of tools/smooth32_bench.cu: one CTA walks the
//                       patch in four slabs of eight z planes (64 KB tile)
//                         phase A  per z slab: f -> tile, subtract (2/h^2) gamma on the boundary cells,
//                                  DST-II along x and y, result -> a CTA-private 256 KB scratch block
//                         phase B  tridiagonal solve along z, register-only, pencils in place in the scratch block
//                         phase C  per z slab: DST-III along y and x, u and its boundary slices -> global
//                       (the scratch block is reused for every patch of the persistent CTA, so it lives in L2)
//   apply3d32_kernel    operator / residual with fused ghost fill, one (patch, z slab) item at a time
//   face_residual_restrict_big_kernel   residual + restriction from face data (see kernels.cuh)
// Reference functions replaced: same as smooth_kernel / apply_kernel / face_residual_restrict_kernel.
#pragma once

namespace tgpu
{
// interface value gamma of entry m on side s (0 on sides without a neighbour); same expressions as
// gamma_all_sides / iface_gamma
template <int D, int N, int MODE>
__device__ __forceinline__ double gamma_entry(const PatchMeta &pm, int p, int s, int m, const FaceVals<D, N, MODE> &fv)
{
	const int ty = pm.nbr_type[s];
	if (ty == NBR_NONE) return 0.0;
	if (ty == NBR_NORMAL) {
		const double a = fv.get(p, pm.parent_idx, pm.orth_on_parent, s, m);
		const double b = fv.get(pm.nbr_idx[s][0], pm.nbr_parent[s], pm.nbr_orth[s], s ^ 1, m);
		return 0.5 * a + 0.5 * b;
	}
	return iface_gamma<D, N, MODE>(pm, p, s, m, fv);
}

constexpr int    S32_ROW = 34, S32_PL = 32 * S32_ROW, S32_TILE = 8 * S32_PL; // slab tile [8][32][34]
constexpr size_t smooth3d32_smem_bytes() { return sizeof(double) * S32_TILE; }
constexpr int    S32_CTAS_PER_SM = 2;

template <bool ZERO_GUESS, bool EMIT, bool PROLONG, bool WRITE_U>
__global__ void __launch_bounds__(TGPU_THREADS, S32_CTAS_PER_SM)
smooth3d32_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ f, double *__restrict__ u,
                  const double *__restrict__ Fin, double *__restrict__ Fout, const double *__restrict__ eig,
                  const double *__restrict__ uc, double *__restrict__ scratch)
{
	constexpr int N = 32, ROW = S32_ROW, PL = S32_PL, M = N * N, NC = N * N * N, SLAB = 8 * M;
	extern __shared__ __align__(16) double S[];
	const int t = threadIdx.x, lo = t & 31, hi = t >> 5; // hi: plane within the slab
	double *  W = scratch + (size_t) blockIdx.x * NC;     // CTA-private block, [z][k_y][k_x]
	Mags<N>   mg;
	mg.load();
	pdl_launch_dependents();
	pdl_wait();
	for (int g = blockIdx.x; g < P - p0; g += gridDim.x) {
		const int        p    = p0 + g;
		const PatchMeta &pm   = meta[p];
		const double     h2   = pm.h2;
		const double     cfac = 2.0 * pm.inv_h2;
		const FaceVals<3, N, PROLONG ? FV_PROLONG : FV_PLAIN> fv{Fin, uc, meta};
		double v[N];
		// ---------------- phase A ----------------
		for (int zs = 0; zs < 4; zs++) {
			const double *src = f + (size_t) p * NC + zs * SLAB;
#pragma unroll
			for (int i = 0; i < 16; i++) {
				const int c = t + TGPU_THREADS * i, row = c >> 4; // 16-byte chunk c of the slab, row = y + 32 z_l
				cp_async16(S + (row & 31) * ROW + (row >> 5) * PL + (c & 15) * 2, src + c * 2);
			}
			cp_async_commit();
			double gx0 = 0.0, gx1 = 0.0;
			if (!ZERO_GUESS) {
				const int m = lo + N * (zs * 8 + hi); // x faces: entry (y, z); y faces: entry (x, z)
				gx0         = cfac * gamma_entry(pm, p, 0, m, fv);
				gx1         = cfac * gamma_entry(pm, p, 1, m, fv);
				const double gy0 = cfac * gamma_entry(pm, p, 2, m, fv);
				const double gy1 = cfac * gamma_entry(pm, p, 3, m, fv);
				cp_async_wait<0>();
				__syncthreads();
				S[lo + hi * PL] -= gy0;
				S[lo + (N - 1) * ROW + hi * PL] -= gy1;
				__syncthreads();
				if (zs == 0 || zs == 3) { // z faces: 1024 entries (x, y), four per thread
					const int s = zs == 0 ? 4 : 5, zl = zs == 0 ? 0 : 7;
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const int mm = t + TGPU_THREADS * i;
						S[(mm & 31) + (mm >> 5) * ROW + zl * PL] -= cfac * gamma_entry(pm, p, s, mm, fv);
					}
					__syncthreads();
				}
			} else {
				cp_async_wait<0>();
				__syncthreads();
			}
			double2 *rowp = reinterpret_cast<double2 *>(S + lo * ROW + hi * PL); // row (y, z_l) = (lo, hi)
#pragma unroll
			for (int j = 0; j < N / 2; j++) {
				const double2 d = rowp[j];
				v[2 * j]        = d.x;
				v[2 * j + 1]    = d.y;
			}
			if (!ZERO_GUESS) {
				v[0] -= gx0;
				v[N - 1] -= gx1;
			}
			Dst2<N, N>::run(v, mg);
#pragma unroll
			for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
			__syncthreads();
			double *q = S + lo + hi * PL; // y pencil (x, z_l) = (lo, hi)
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = q[k * ROW];
			Dst2<N, N>::run(v, mg);
			double *w = W + (size_t) (zs * 8 + hi) * M + lo;
#pragma unroll
			for (int k = 0; k < N; k++) w[k * N] = v[k];
			__syncthreads(); // the tile is refilled next
		}
		// ---------------- phase B: z pencils (k_x, k_y) = (lo, 8 ys + hi), in place in W ----------------
		for (int ys = 0; ys < 4; ys++) {
			const int ky = ys * 8 + hi;
			double *  q  = W + ky * N + lo;
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = q[(size_t) k * M];
			// x and y are diagonalised: a tridiagonal system along z per (k_x, k_y) (TriSolve, smooth3d16.cuh);
			// eig = the table of elimination multipliers [17][k_y][k_x]
			TriSolve<N, M>::forward(v, eig + ky * N + lo, h2 * (4.0 / (N * N)));
			TriSolve<N, M>::backward(v, eig + ky * N + lo);
#pragma unroll
			for (int k = 0; k < N; k++) q[(size_t) k * M] = v[k];
		}
		__syncthreads();
		// ---------------- phase C ----------------
		double *Fp = EMIT ? Fout + (size_t) p * 6 * M : nullptr;
		for (int zs = 0; zs < 4; zs++) {
			const int     z = zs * 8 + hi;
			const double *w = W + (size_t) z * M + lo; // y pencil (x, z) = (lo, z)
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = w[k * N];
			Dst3<N, N>::run(v, mg);
			double *q = S + lo + hi * PL;
#pragma unroll
			for (int k = 0; k < N; k++) q[k * ROW] = v[k];
			__syncthreads();
			double2 *rowp = reinterpret_cast<double2 *>(S + lo * ROW + hi * PL); // row (y, z_l) = (lo, hi)
#pragma unroll
			for (int j = 0; j < N / 2; j++) {
				const double2 d = rowp[j];
				v[2 * j]        = d.x;
				v[2 * j + 1]    = d.y;
			}
			Dst3<N, N>::run(v, mg);
#pragma unroll
			for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
			if (EMIT) { // x faces: entry (y, z)
				Fp[0 * M + lo + N * z] = v[0];
				Fp[1 * M + lo + N * z] = v[N - 1];
			}
			__syncthreads();
			if (WRITE_U) {
				double *dst = u + (size_t) p * NC + zs * SLAB;
#pragma unroll
				for (int i = 0; i < 16; i++) {
					const int c = t + TGPU_THREADS * i, row = c >> 4;
					*reinterpret_cast<double2 *>(dst + c * 2) =
					*reinterpret_cast<const double2 *>(S + (row & 31) * ROW + (row >> 5) * PL + (c & 15) * 2);
				}
			}
			if (EMIT) {
				Fp[2 * M + lo + N * z] = S[lo + hi * PL]; // y faces: entry (x, z)
				Fp[3 * M + lo + N * z] = S[lo + (N - 1) * ROW + hi * PL];
				if (zs == 0 || zs == 3) { // z faces: entry (x, y)
					const int s = zs == 0 ? 4 : 5, zl = zs == 0 ? 0 : 7;
#pragma unroll
					for (int i = 0; i < 4; i++) {
						const int mm    = t + TGPU_THREADS * i;
						Fp[s * M + mm] = S[(mm & 31) + (mm >> 5) * ROW + zl * PL];
					}
				}
			}
			__syncthreads();
		}
	}
}

// ---------------------------------------------------------------------------------------------
// smooth3d32c_kernel: the 32^3 patch solve on a pair of SMs (thread-block cluster of two CTAs), whole patch
// resident in shared memory, no scratch block.
//   The z axis is not transformed (TriSolve, kernels.cuh): with x and y diagonalised, every (k_x, k_y) pencil
//   is a tridiagonal system along z whose two-sided elimination runs from z = 0 upwards and from z = 31
//   downwards with the same multipliers.  That is exactly a split over two CTAs: CTA 0 owns the planes
//   z = 0..15, CTA 1 the planes z = 31..16 (local plane j <-> elimination step j in both), and the only data the
//   halves exchange is the last plane of eliminated right-hand sides (8 KB each way, read through distributed
//   shared memory after one cluster barrier) for the 2 x 2 system in the middle.
//   Per CTA: 512 threads = 16 warps, warp w owns local plane w for the y and x transforms (warp-local
//   transposes), the z recurrences are element-wise over the planes (two pencils per thread).
//   f goes straight from memory into the y pencils (a warp's load = one 256-byte row) and u straight from the
//   y pencils back to memory: HBM sees f once and u once, shared memory holds the 16 planes (136 KB).
// Same arithmetic as smooth3d32_kernel / smooth_kernel (SchurHelper.h:319-331, FftwPatchSolver.h:174-206).
// ---------------------------------------------------------------------------------------------
constexpr int    C32_THREADS = 512, C32_ROW = 34, C32_PL = 32 * C32_ROW, C32_TILE = 16 * C32_PL;
// tile + two exchange planes (double buffered) + z-face interface values + per-warp x/y-face values + x-face output staging
#ifndef C32_ZREG
#define C32_ZREG 0 // 1: z pencils stay in registers across the cluster barrier instead of a store / reload of the eliminated planes
                   // (measured on 4096 patches: -3 % on the variants without gathers, +9 % (spills) on the two the fused cycle uses)
#endif
#ifndef C32_FACETAIL
#define C32_FACETAIL 1 // faces-only sweeps from a zero guess: dot products + 60 one-dimensional transforms instead of full inverse transforms
#endif
constexpr int    C32_FV = 65; // row pitch of the face-vector staging [32][65] (faces-only sweeps, see the tail of the kernel)
// (188 KB: the next shared-memory carve-out step, 228 KB, would leave the L1 28 KB instead of 60 KB - measured 4-7 % slower)
constexpr size_t smooth3d32c_smem_bytes() { return sizeof(double) * (C32_TILE + 2 * 1024 + 1024 + 16 * 128 + 16 * 64); }
static_assert(32 * C32_FV <= 1024 + 16 * 128, "the face-vector staging aliases the interface-value buffers of sweeps from a zero guess");
__device__ __forceinline__ unsigned cluster_ctarank()
{
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// value at the same shared-memory address in CTA `rank` of the cluster
__device__ __forceinline__ double ld_dsmem(const double *local, unsigned rank)
{
	unsigned a = (unsigned) __cvta_generic_to_shared(local), ra;
	double   v;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(a), "r"(rank));
	asm volatile("ld.shared::cluster.f64 %0, [%1];\n" : "=d"(v) : "r"(ra) : "memory");
	return v;
}
template <bool PROLONG>
__device__ __noinline__ double gamma_entry32(const PatchMeta *__restrict__ meta, int p, int s, int m, const double *__restrict__ F,
                                             const double *__restrict__ uc)
{
	const FaceVals<3, 32, PROLONG ? FV_PROLONG : FV_PLAIN> fv{F, uc, meta};
	return gamma_entry(meta[p], p, s, m, fv);
}
// One interface value split into "issue the loads" and "combine": same-level neighbours inline (same expressions
// as gamma_entry), anything else through the general code at combine time.
template <bool PROLONG> struct Gam32 {
	double a0, b0, a1, b1;
	int    slow;
	__device__ __forceinline__ void issue(const PatchMeta &pm, int p, int s, int m, const double *__restrict__ F, const double *__restrict__ uc)
	{
		constexpr int N = 32, M = N * N, NC = M * N;
		const int     ty = pm.nbr_type[s];
		slow             = ty > NBR_NORMAL;
		a0 = b0 = a1 = b1 = 0.0;
		if (ty == NBR_NORMAL) {
			a0 = __ldg(F + ((size_t) p * 6 + s) * M + m);
			b0 = __ldg(F + ((size_t) pm.nbr_idx[s][0] * 6 + (s ^ 1)) * M + m);
			if (PROLONG) {
				int c[3];
				if (pm.parent_idx >= 0) {
					face_cell<3, N>(s, m, c);
					a1 = __ldg(uc + (size_t) pm.parent_idx * NC + parent_cell<3, N>(pm.orth_on_parent, c));
				}
				const int qp = pm.nbr_parent[s];
				if (qp >= 0) { // < 0: halo slot whose face arrived with the correction added
					face_cell<3, N>(s ^ 1, m, c);
					b1 = __ldg(uc + (size_t) qp * NC + parent_cell<3, N>(pm.nbr_orth[s], c));
				}
			}
		}
	}
	__device__ __forceinline__ double finish(const PatchMeta *__restrict__ meta, int p, int s, int m, const double *__restrict__ F,
	                                         const double *__restrict__ uc, double cfac) const
	{
		if (slow) return cfac * gamma_entry32<PROLONG>(meta, p, s, m, F, uc);
		return cfac * (0.5 * (a0 + a1) + 0.5 * (b0 + b1));
	}
};

// Gather descriptor of one patch side (the 32^3 analogue of GDesc16, smooth3d16.cuh): everything that is uniform over the
// 1024 face entries, resolved once per patch by one thread, so that an entry's four loads are "base + per-thread constant":
//   (2/h^2) gamma = c (0.5 (a0[m] + a1[o]) + 0.5 (b0[m] + wb b1[o])),  o = the entry's offset on the parent's plane
// a0 / b0: own / neighbour face slice, a1 / b1: the parents' cells under them (DrctIntp.h:92-111, only with the prolongation
// fused in), c = 0 on sides without a neighbour, wb = 0 for halo faces that already carry the correction, SA / SB: the
// parent is a refined patch (two entries per coarse cell) or the same patch one level up (copy-add).  Sides with coarse /
// fine neighbours are flagged GD_SLOW and take gamma_entry32.
#ifndef C32_GDESC
#define C32_GDESC 1 // 0: per-thread neighbour-table reads and index arithmetic for every entry (Gam32)
#endif
struct __align__(16) GDesc32 {
	unsigned a0, b0, a1, b1; // element offsets into F (a0, b0) and into uc (a1, b1)
	double   c;
	int      flags, pad;
};
struct GPatch32 {
	GDesc32 d[6];
	int     neu, pad[3]; // Neumann bits of the patch (levels with Neumann patches: the cluster kernel skips them)
};
template <bool PROLONG>
__device__ __forceinline__ void make_gdesc32(const PatchMeta &pm, int p, int s, GPatch32 &out)
{
	GDesc32 & d  = out.d[s];
	const int ty = pm.nbr_type[s];
	const int ax = s >> 1;
	const int st = (ax == 0) ? 1 : (ax == 1 ? 32 : 1024); // stride of the face-normal axis
	d.a0 = d.b0 = ((unsigned) p * 6 + s) * 1024;
	d.a1 = d.b1 = 0;
	d.c         = (ty == NBR_NONE) ? 0.0 : 2.0 * pm.inv_h2;
	int fl      = GD_SA | GD_SB | GD_WB;
	if (ty == NBR_NORMAL) {
		d.b0 = ((unsigned) pm.nbr_idx[s][0] * 6 + (s ^ 1)) * 1024;
		if (PROLONG) {
			const int o = pm.orth_on_parent, qp = pm.nbr_parent[s], qo = pm.nbr_orth[s];
			if (o >= 0) d.a1 = (unsigned) pm.parent_idx * 32768 + 16 * ((o & 1) + 32 * ((o >> 1) & 1) + 1024 * ((o >> 2) & 1)) + ((s & 1) ? 15 * st : 0);
			else d.a1 = (unsigned) pm.parent_idx * 32768 + ((s & 1) ? 31 * st : 0), fl &= ~GD_SA;
			if (qp < 0) d.b1 = d.a1, fl = (fl & ~(GD_SB | GD_WB)) | ((fl & GD_SA) ? GD_SB : 0);
			else if (qo >= 0) d.b1 = (unsigned) qp * 32768 + 16 * ((qo & 1) + 32 * ((qo >> 1) & 1) + 1024 * ((qo >> 2) & 1)) + ((s & 1) ? 0 : 15 * st);
			else d.b1 = (unsigned) qp * 32768 + ((s & 1) ? 0 : 31 * st), fl &= ~GD_SB;
		}
	} else if (ty > NBR_NORMAL) {
		fl = GD_SA | GD_SB | GD_WB | GD_SLOW; // harmless addresses; the value comes from gamma_entry32
	}
	d.flags = fl;
}
template <bool PROLONG> struct SideGamma32 {
	double a0, a1, b0, b1;
	// AX: face-normal axis; entry m = lo + 32 hi lies over cell (lo >> s) * A + (hi >> s) * B of the parent's plane
	template <int AX>
	__device__ __forceinline__ void issue(const GDesc32 &d, int m, int lo, int hi, const double *__restrict__ F, const double *__restrict__ uc)
	{
		constexpr int A = (AX == 0) ? 32 : 1, B = (AX == 2) ? 32 : 1024;
		const uint4   o = *reinterpret_cast<const uint4 *>(&d);
		a0 = __ldg(F + o.x + m);
		b0 = __ldg(F + o.y + m);
		a1 = b1 = 0.0;
		if (PROLONG) {
			const int fl = d.flags, sa = fl & GD_SA, sb = (fl >> 1) & 1;
			a1 = __ldg(uc + o.z + ((lo >> sa) * A + (hi >> sa) * B));
			b1 = __ldg(uc + o.w + ((lo >> sb) * A + (hi >> sb) * B));
		}
	}
	__device__ __forceinline__ double finish(const GDesc32 &d, const PatchMeta *__restrict__ meta, int p, int s, int m, const double *__restrict__ F,
	                                         const double *__restrict__ uc) const
	{
		const double c  = d.c;
		const int    fl = d.flags;
		double       g  = c * (0.5 * (a0 + a1) + 0.5 * (b0 + ((fl & GD_WB) ? b1 : 0.0)));
		if (fl & GD_SLOW) g = c * gamma_entry32<PROLONG>(meta, p, s, m, F, uc);
		return g;
	}
};

// HALO = false compiles the multi-GPU hand-over (HaloSync / halo_push_cta) out.  The host launches HALO = !ZERO_GUESS on
// any number of GPUs: measured on one GPU (config D, same box, alternating runs) the post-sweep instantiation WITH the
// hand-over code is 1 % faster than the one without (7.05 vs 7.13 ms; register allocation at the 128-register cap).
template <bool ZERO_GUESS, bool EMIT, bool PROLONG, bool WRITE_U, bool HALO = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C32_THREADS, 1)
smooth3d32c_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ f, double *__restrict__ u,
                   const double *__restrict__ Fin, double *__restrict__ Fout, const double *__restrict__ tri,
                   const double *__restrict__ uc, HaloSync hs = HaloSync{}, int skip_neumann = 0)
{
	// skip_neumann != 0: patches with Neumann domain sides are left to smooth3d32n_kernel (launched over the same range
	// with only_neumann); both CTAs of the cluster skip together and keep gathering the next patch's interface values
	constexpr int N = 32, ROW = C32_ROW, PL = C32_PL, M = N * N, NC = N * N * N;
	extern __shared__ __align__(16) double S[];
	double *       X    = S + C32_TILE;  // [2][M] last eliminated plane, for the peer
	double *       GZ   = X + 2 * M;     // [M] (2/h^2) gamma on this half's z face
	double *       GXY  = GZ + M;        // [16 warps][4][32] x- and y-face values of a plane
	double *       EX   = GXY + 16 * 128; // [16 warps][2][32] staging of the new x-face slices
	double *       FV   = GZ;             // [32][C32_FV] faces-only sweeps from a zero guess (GZ and GXY are unused there): partially
	                                      // transformed face vectors, entry j of vector i at j * C32_FV + i
	const int      t = threadIdx.x, lane = t & 31, w = t >> 5;
	const unsigned rank = cluster_ctarank();
	const int      z    = rank == 0 ? w : N - 1 - w; // plane of the patch behind local plane w
	const int      sz   = rank == 0 ? 4 : 5;         // the z side this half touches
	const int      ncl = gridDim.x / 2, npatch = P - p0;
	double *       gxy  = GXY + w * 128;
	const int      mf   = lane + N * z; // x faces: entry (y, z) = (lane, z); y faces: entry (x, z) = (lane, z)
	Mags<N>        mg;
	mg.load();
	pdl_launch_dependents();
	pdl_wait();
	if (!ZERO_GUESS) if (HALO) halo_push<3, 32>(hs, meta);
	int g = blockIdx.x / 2;
	// multi-GPU: each CTA polls the peers' flags itself before the first patch whose gamma needs halo faces (HaloSync)
	bool halo_ok = false;
#if C32_GDESC
	// neighbour-table entry of the patch after the next (staged with cp.async) and the gather descriptors of the next
	// patch (GD[(it + 1) & 1]) / the one after it, resolved by lane 31 of warps 0-5
	constexpr int                   MW = (int) (sizeof(PatchMeta) / sizeof(double));
	__shared__ __align__(16) double metaS[MW];
	__shared__ GPatch32             GD[2];
	auto describe = [&](const PatchMeta &pm, int q, int slot) {
		if (lane == 31 && w < 6) make_gdesc32<PROLONG>(pm, q, w, GD[slot]);
		if (lane == 31 && w == 6) GD[slot].neu = pm.neumann;
	};
	const int zlo = t & 31, zhi = t >> 5; // z-face entries t and t + 512: (x, y) = (zlo, zhi), (zlo, zhi + 16)
	if (!ZERO_GUESS && g < npatch) { // first patch of this cluster: nothing to hide the gathers behind
		const int p = p0 + g;
		if (HALO) halo_wait_cta(hs, p, halo_ok);
		describe(meta[p], p, 0);
		if (g + ncl < npatch) describe(meta[p + ncl], p + ncl, 1);
		__syncthreads();
		// all loads of the six interface values in flight at once: one memory round trip (coarse levels: one patch per cluster)
		SideGamma32<PROLONG> g6[6];
		g6[0].template issue<0>(GD[0].d[0], mf, lane, z, Fin, uc);
		g6[1].template issue<0>(GD[0].d[1], mf, lane, z, Fin, uc);
		g6[2].template issue<1>(GD[0].d[2], mf, lane, z, Fin, uc);
		g6[3].template issue<1>(GD[0].d[3], mf, lane, z, Fin, uc);
		g6[4].template issue<2>(GD[0].d[sz], t, zlo, zhi, Fin, uc);
		g6[5].template issue<2>(GD[0].d[sz], t + 512, zlo, zhi + 16, Fin, uc);
#pragma unroll
		for (int s = 0; s < 4; s++) gxy[s * 32 + lane] = g6[s].finish(GD[0].d[s], meta, p, s, mf, Fin, uc);
		GZ[t]       = g6[4].finish(GD[0].d[sz], meta, p, sz, t, Fin, uc);
		GZ[t + 512] = g6[5].finish(GD[0].d[sz], meta, p, sz, t + 512, Fin, uc);
		__syncthreads();
	}
#else
	if (!ZERO_GUESS && g < npatch) { // first patch of this cluster: nothing to hide the gathers behind
		const int    p    = p0 + g;
		if (HALO) halo_wait_cta(hs, p, halo_ok);
		const PatchMeta &pq   = meta[p];
		const double     cfac = 2.0 * pq.inv_h2;
		// all loads of the six interface values in flight at once: one memory round trip (coarse levels: one patch per cluster)
		Gam32<PROLONG> g6[6];
#pragma unroll
		for (int s = 0; s < 4; s++) g6[s].issue(pq, p, s, mf, Fin, uc);
		g6[4].issue(pq, p, sz, t, Fin, uc);
		g6[5].issue(pq, p, sz, t + 512, Fin, uc);
#pragma unroll
		for (int s = 0; s < 4; s++) gxy[s * 32 + lane] = g6[s].finish(meta, p, s, mf, Fin, uc, cfac);
		GZ[t]       = g6[4].finish(meta, p, sz, t, Fin, uc, cfac);
		GZ[t + 512] = g6[5].finish(meta, p, sz, t + 512, Fin, uc, cfac);
		__syncthreads();
	}
#endif
	for (int it = 0; g < npatch; g += ncl, it++) {
		const int    p    = p0 + g;
		const bool   next = g + ncl < npatch;
		const int    pn   = p + ncl;
		const double h2   = meta[p].h2;
		double       v[N];
#if C32_GDESC
		// table entry of the patch after the next -> metaS (described after this iteration's first barrier)
		if (!ZERO_GUESS && t < MW && g + 2 * ncl < npatch) cp_async8(&metaS[t], reinterpret_cast<const double *>(meta + pn + ncl) + t, true);
		if (skip_neumann && (ZERO_GUESS ? meta[p].neumann : GD[it & 1].neu)) { // cluster-uniform: this patch belongs to the general path
			if (!ZERO_GUESS) {
				// keep the pipeline of the next patch's interface values going: all six at once, nothing to hide behind
				cp_async_wait_all();
				__syncthreads(); // metaS has landed; gxy / GZ of this (skipped) patch are free
				const GPatch32 &gq = GD[(it + 1) & 1];
				if (next) {
					if (HALO) halo_wait_cta(hs, pn, halo_ok);
					SideGamma32<PROLONG> s6[6];
					s6[0].template issue<0>(gq.d[0], mf, lane, z, Fin, uc);
					s6[1].template issue<0>(gq.d[1], mf, lane, z, Fin, uc);
					s6[2].template issue<1>(gq.d[2], mf, lane, z, Fin, uc);
					s6[3].template issue<1>(gq.d[3], mf, lane, z, Fin, uc);
					s6[4].template issue<2>(gq.d[sz], t, zlo, zhi, Fin, uc);
					s6[5].template issue<2>(gq.d[sz], t + 512, zlo, zhi + 16, Fin, uc);
#pragma unroll
					for (int s = 0; s < 4; s++) gxy[s * 32 + lane] = s6[s].finish(gq.d[s], meta, pn, s, mf, Fin, uc);
					GZ[t]       = s6[4].finish(gq.d[sz], meta, pn, sz, t, Fin, uc);
					GZ[t + 512] = s6[5].finish(gq.d[sz], meta, pn, sz, t + 512, Fin, uc);
					if (g + 2 * ncl < npatch) describe(*reinterpret_cast<const PatchMeta *>(metaS), pn + ncl, it & 1);
				}
				__syncthreads(); // the values and descriptors are visible to the next iteration
			}
			cluster_sync_all(); // keeps the pair in step: the exchange planes alternate by iteration parity
			continue;
		}
#endif
		{ // y forward: pencil (x, z) = (lane, z), straight from memory
			const double *fp = f + (size_t) p * NC + (size_t) z * M + lane;
#pragma unroll
			for (int k = 0; k < N; k++) v[k] = __ldcs(fp + k * N);
			if (next) { // the same plane of the next patch -> L2 (64 lines)
				const double *fn = f + (size_t) pn * NC + (size_t) z * M + lane * 32;
				prefetch_l2(fn);
				prefetch_l2(fn + 16);
			}
			if (!ZERO_GUESS) {
				v[0] -= gxy[64 + lane];
				v[N - 1] -= gxy[96 + lane];
				if (lane == 0 || lane == N - 1) {
					const double *q = gxy + (lane ? 32 : 0);
#pragma unroll
					for (int k = 0; k < N; k++) v[k] -= q[k];
				}
				if (w == 0) {
#pragma unroll
					for (int k = 0; k < N; k++) v[k] -= GZ[lane + N * k];
				}
			}
			Dst2<N, N>::run(v, mg);
			double *col = S + w * PL + lane;
#pragma unroll
			for (int k = 0; k < N; k++) col[k * ROW] = v[k];
		}
		__syncwarp();
		double2 *rowp = reinterpret_cast<double2 *>(S + w * PL + lane * ROW); // row (k_y, plane) = (lane, w)
		{ // x forward
#pragma unroll
			for (int j = 0; j < N / 2; j++) {
				const double2 d = rowp[j];
				v[2 * j]        = d.x;
				v[2 * j + 1]    = d.y;
			}
			Dst2<N, N>::run(v, mg);
#pragma unroll
			for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
		}
#if C32_GDESC
		if (!ZERO_GUESS) cp_async_wait_all(); // metaS has landed (made visible by the barrier)
#endif
		__syncthreads();
		if (HALO && !ZERO_GUESS && next) halo_wait_cta(hs, pn, halo_ok); // gamma of patch pn is gathered from here on
		// z: elimination step j on local plane j, in place; pencils (k_x, k_y) = (lane, w) and (lane, w + 16).
		// The interface values of the NEXT patch are gathered around this phase (the transform registers are free here).
		// (two batches of three: loads issued before the elimination / the back substitution, combined after it)
#if C32_GDESC
#ifndef C32_GATHER6
#define C32_GATHER6 0 // 1: all six sides' loads in flight across the whole z phase instead of two batches of three (measured: -5 % on the sweeps without prolongation, none with it, and spills)
#endif
		SideGamma32<PROLONG> gm[C32_GATHER6 ? 6 : 3];
		const GPatch32 &     gp = GD[(it + 1) & 1]; // descriptors of patch pn
		if (!ZERO_GUESS && next) {
			gm[0].template issue<0>(gp.d[0], mf, lane, z, Fin, uc);
			gm[1].template issue<0>(gp.d[1], mf, lane, z, Fin, uc);
			gm[2].template issue<1>(gp.d[2], mf, lane, z, Fin, uc);
#if C32_GATHER6
			gm[3].template issue<1>(gp.d[3], mf, lane, z, Fin, uc);
			gm[4].template issue<2>(gp.d[sz], t, zlo, zhi, Fin, uc);
			gm[5].template issue<2>(gp.d[sz], t + 512, zlo, zhi + 16, Fin, uc);
#endif
			// descriptors of the patch after the next; GD[it & 1] (patch p) was last read during the previous iteration
			if (g + 2 * ncl < npatch) describe(*reinterpret_cast<const PatchMeta *>(metaS), pn + ncl, it & 1);
		}
#else
		Gam32<PROLONG> gm[3];
		double         cfn = 0.0;
		if (!ZERO_GUESS && next) {
			const PatchMeta &pq = meta[pn];
			cfn                 = 2.0 * pq.inv_h2;
#pragma unroll
			for (int s = 0; s < 3; s++) gm[s].issue(pq, pn, s, mf, Fin, uc);
		}
#endif
		const double hsc = h2 * (4.0 / (N * N));
		double *     Xo = X + (it & 1) * M;
#if C32_ZREG
		double       rz[2][16]; // the two pencils stay in registers across the cluster barrier (no store / reload of the eliminated planes)
#endif
#pragma unroll
		for (int q = 0; q < 2; q++) {
			const int     ky = w + 16 * q;
			const double *zp = S + ky * ROW + lane;
			const double *tb = tri + ky * N + lane;
			double        rho = 0.0;
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const double a = __ldg(tb + j * M);
				const double r = zp[j * PL] * (hsc * a);
				rho            = (j == 0) ? r : fma(-a, rho, r);
#if C32_ZREG
				rz[q][j]       = rho;
#else
				const_cast<double *>(zp)[j * PL] = rho;
#endif
			}
			Xo[ky * N + lane] = rho;
		}
		if (!ZERO_GUESS && next) { // (the buffers were consumed before this iteration's first barrier)
#if C32_GDESC && C32_GATHER6
			// (nothing here: the six values are combined after the back substitution)
#elif C32_GDESC
#pragma unroll
			for (int s = 0; s < 3; s++) gxy[s * 32 + lane] = gm[s].finish(gp.d[s], meta, pn, s, mf, Fin, uc);
			gm[0].template issue<1>(gp.d[3], mf, lane, z, Fin, uc);
			gm[1].template issue<2>(gp.d[sz], t, zlo, zhi, Fin, uc);
			gm[2].template issue<2>(gp.d[sz], t + 512, zlo, zhi + 16, Fin, uc);
#else
#pragma unroll
			for (int s = 0; s < 3; s++) gxy[s * 32 + lane] = gm[s].finish(meta, pn, s, mf, Fin, uc, cfn);
			const PatchMeta &pq = meta[pn];
			gm[0].issue(pq, pn, 3, mf, Fin, uc);
			gm[1].issue(pq, pn, sz, t, Fin, uc);
			gm[2].issue(pq, pn, sz, t + 512, Fin, uc);
#endif
		}
		cluster_sync_all(); // both halves are eliminated; the peer's last plane is readable
#pragma unroll
		for (int q = 0; q < 2; q++) {
			const int     ky = w + 16 * q;
			double *      zp = S + ky * ROW + lane;
			const double *tb = tri + ky * N + lane;
			const double  a15 = __ldg(tb + 15 * M), kap = __ldg(tb + 16 * M);
#if C32_ZREG
			double        y   = kap * fma(-a15, ld_dsmem(Xo + ky * N + lane, rank ^ 1), rz[q][15]);
#else
			double        y   = kap * fma(-a15, ld_dsmem(Xo + ky * N + lane, rank ^ 1), zp[15 * PL]);
#endif
			zp[15 * PL]       = y;
#pragma unroll
			for (int j = 14; j >= 0; j--) {
#if C32_ZREG
				y          = fma(-__ldg(tb + j * M), y, rz[q][j]);
#else
				y          = fma(-__ldg(tb + j * M), y, zp[j * PL]);
#endif
				zp[j * PL] = y;
			}
		}
		if (!ZERO_GUESS && next) {
#if C32_GDESC && C32_GATHER6
#pragma unroll
			for (int s = 0; s < 4; s++) gxy[s * 32 + lane] = gm[s].finish(gp.d[s], meta, pn, s, mf, Fin, uc);
			GZ[t]       = gm[4].finish(gp.d[sz], meta, pn, sz, t, Fin, uc);
			GZ[t + 512] = gm[5].finish(gp.d[sz], meta, pn, sz, t + 512, Fin, uc);
#elif C32_GDESC
			gxy[96 + lane] = gm[0].finish(gp.d[3], meta, pn, 3, mf, Fin, uc);
			GZ[t]          = gm[1].finish(gp.d[sz], meta, pn, sz, t, Fin, uc);
			GZ[t + 512]    = gm[2].finish(gp.d[sz], meta, pn, sz, t + 512, Fin, uc);
#else
			gxy[96 + lane] = gm[0].finish(meta, pn, 3, mf, Fin, uc, cfn);
			GZ[t]          = gm[1].finish(meta, pn, sz, t, Fin, uc, cfn);
			GZ[t + 512]    = gm[2].finish(meta, pn, sz, t + 512, Fin, uc, cfn);
#endif
		}
		__syncthreads();
		constexpr bool FACE_TAIL = C32_FACETAIL && ZERO_GUESS && !WRITE_U;
		if (!FACE_TAIL || w == 0) {
			// full inverse transforms: every plane of a sweep that writes u; in a faces-only sweep only the plane that is
			// this half's z face (local plane 0)
			{ // x inverse
#pragma unroll
				for (int j = 0; j < N / 2; j++) {
					const double2 d = rowp[j];
					v[2 * j]        = d.x;
					v[2 * j + 1]    = d.y;
				}
				Dst3<N, N>::run(v, mg);
#pragma unroll
				for (int j = 0; j < N / 2; j++) rowp[j] = make_double2(v[2 * j], v[2 * j + 1]);
			}
			__syncwarp();
			{ // y inverse: pencil (x, z) = (lane, z), straight to memory
				const double *col = S + w * PL + lane;
#pragma unroll
				for (int k = 0; k < N; k++) v[k] = col[k * ROW];
				Dst3<N, N>::run(v, mg);
				if (WRITE_U) {
					double *up = u + (size_t) p * NC + (size_t) z * M + lane;
#pragma unroll
					for (int k = 0; k < N; k++) __stcs(up + k * N, v[k]);
				}
				if (EMIT) {
					double *Fp = Fout + (size_t) p * 6 * M;
					Fp[2 * M + mf] = v[0]; // y faces: entry (x, z)
					Fp[3 * M + mf] = v[N - 1];
					if (w == 0) { // z face of this half: entries (x, y)
						double *Fz = Fp + sz * M + lane;
#pragma unroll
						for (int k = 0; k < N; k++) Fz[N * k] = v[k];
					}
					double *ex = EX + w * 64;
					if (lane == 0 || lane == N - 1) { // x faces: entries (y, z), held by lanes 0 and 31
						double *q = ex + (lane ? 32 : 0);
#pragma unroll
						for (int k = 0; k < N; k++) q[k] = v[k];
					}
					__syncwarp();
					Fp[0 * M + mf] = ex[lane];
					Fp[1 * M + mf] = ex[32 + lane];
				}
			}
		} else {
			// Faces-only sweep, planes other than the z face: only u(0 | 31, y, z) and u(x, 0 | 31, z) are wanted.  With
			// T = the DST-III matrix (DftPatchSolver.h:269-281), T[0][j] = sin(pi (j+1) / 2n), T[0][n-1] = 1/2 and
			// T[n-1][j] = (-1)^j T[0][j], contract the axis normal to the face first - two dot products per row / column of
			// the plane instead of a transform - and leave the 4 x 15 remaining 1-D transforms of the CTA to two warps.
			double E = 0.0, O = 0.0;
#pragma unroll
			for (int j = 0; j < N / 2; j++) { // row (k_y, plane) = (lane, w): contract k_x
				const double2 d = rowp[j];
				E               = fma(mg.sinq(2 * j + 1), d.x, E);
				O               = fma((2 * j + 1 == N - 1) ? 0.5 : mg.sinq(2 * j + 2), d.y, O);
			}
			const int vi = 4 * (w - 1); // vectors 4 (w - 1) + {0: x = 0, 1: x = 31, 2: y = 0, 3: y = 31}, entry index = lane
			FV[lane * C32_FV + vi]     = E + O;
			FV[lane * C32_FV + vi + 1] = E - O;
			const double *col = S + w * PL + lane; // column (k_x, plane) = (lane, w): contract k_y
			E = O = 0.0;
#pragma unroll
			for (int k = 0; k < N; k += 2) {
				E = fma(mg.sinq(k + 1), col[k * ROW], E);
				O = fma((k + 1 == N - 1) ? 0.5 : mg.sinq(k + 2), col[(k + 1) * ROW], O);
			}
			FV[lane * C32_FV + vi + 2] = E + O;
			FV[lane * C32_FV + vi + 3] = E - O;
		}
		if (FACE_TAIL) {
			__syncthreads();
			const int vec = (w - 1) * 32 + lane; // warps 1 and 2: one face vector per lane
			if ((w == 1 || w == 2) && vec < 60) {
#pragma unroll
				for (int j = 0; j < N; j++) v[j] = FV[j * C32_FV + vec];
				Dst3<N, N>::run(v, mg);
				const int lp = 1 + (vec >> 2), kind = vec & 3;             // local plane, which face
				const int zz = rank == 0 ? lp : N - 1 - lp;
				double *  Fq = Fout + (size_t) p * 6 * M + kind * M + N * zz; // x faces: entries (y, z); y faces: entries (x, z)
#pragma unroll
				for (int j = 0; j < N / 2; j++) *reinterpret_cast<double2 *>(Fq + 2 * j) = make_double2(v[2 * j], v[2 * j + 1]);
			}
		}
		__syncwarp();
	}
	cluster_sync_all(); // a CTA must not exit while its peer may still read its exchange plane
	if (!ZERO_GUESS) if (HALO) halo_finish(hs);
}

// ---------------------------------------------------------------------------------------------
// smooth3d32n_kernel: 32^3 levels that contain patches with Neumann domain sides (PatchSolvers/FftwPatchSolver.h:
// 115-127,197; DftPatchSolver.h:115-127,150-165).  General path, same arithmetic as the Neumann branch of
// smooth_kernel: per axis the transform pair and eigenvalues follow the two closures (axis_kind), dense 32 x 32
// transforms with the matrices of DftPatchSolver.h:237-289 (mats), eigenvalue sums formed on the fly (lam), zero mode
// removed on an all-Neumann patch.  One CTA per patch, the patch lives in a CTA-private 256 KB block of `scratch`
// (L2-resident: the block is reused for every patch of the CTA); the variant switches are run-time arguments.
// Correctness path, not a fast one: 6 x 1024 multiply-adds per cell.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TGPU_THREADS)
smooth3d32n_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ f, double *__restrict__ u,
                   const double *__restrict__ Fin, double *__restrict__ Fout, const double *__restrict__ uc,
                   const double *__restrict__ mats, const double *__restrict__ lam, double *__restrict__ scratch, int zero_guess,
                   int emit, int prolong, int write_u, double lam_shift, int only_neumann = 0)
{
	// only_neumann != 0: only patches with Neumann domain sides are swept (the others of the range belong to
	// smooth3d32c_kernel, skip_neumann)
	constexpr int N = 32, M = N * N, NC = M * N;
	const int     t = threadIdx.x;
	double *      W = scratch + (size_t) blockIdx.x * NC;
	Mags<N>       mg;
	mg.load();
	pdl_launch_dependents();
	pdl_wait();
	for (int g = blockIdx.x; g < P - p0; g += gridDim.x) {
		const int        p    = p0 + g;
		const PatchMeta &pm   = meta[p];
		const int        neu  = pm.neumann;
		if (only_neumann && !neu) continue; // CTA-uniform
		const double     cfac = 2.0 * pm.inv_h2, h2 = pm.h2;
		for (int i = t; i < NC; i += TGPU_THREADS) W[i] = __ldg(f + (size_t) p * NC + i);
		__syncthreads();
		if (!zero_guess) { // f - (2/h^2) E^T gamma, side by side (edge cells belong to several sides)
			for (int s = 0; s < 6; s++) {
				if (pm.nbr_type[s] != NBR_NONE) {
					for (int m = t; m < M; m += TGPU_THREADS) {
						const double gm = prolong ? gamma_entry32<true>(meta, p, s, m, Fin, uc) : gamma_entry32<false>(meta, p, s, m, Fin, uc);
						int          c[3];
						face_cell<3, N>(s, m, c);
						W[(c[2] * N + c[1]) * N + c[0]] -= cfac * gm;
					}
				}
				__syncthreads();
			}
		}
		double v[N];
		// forward along z, y, x; then inverse along x, y, z (pass 2 also divides by the eigenvalue sums first)
		for (int pass = 0; pass < 6; pass++) {
			const int      axis = pass < 3 ? 2 - pass : pass - 3;
			const AxisKind ak   = axis_kind(neu, axis);
				const int      step = axis == 0 ? 1 : (axis == 1 ? N : M);
			if (pass == 3) {
				const AxisKind kx = axis_kind(neu, 0), ky = axis_kind(neu, 1), kz = axis_kind(neu, 2);
				const double   scale = h2 * (2.0 / N) * (2.0 / N) * (2.0 / N);
				const bool     singular = neu == 63;
				for (int i = t; i < NC; i += TGPU_THREADS) {
					const double sum = __ldg(lam + kx.lam * N + i % N) + (__ldg(lam + ky.lam * N + (i / N) % N) + __ldg(lam + kz.lam * N + i / M)) + lam_shift * h2;
					W[i]             = (singular && i == 0) ? 0.0 : W[i] * scale / sum;
				}
				__syncthreads();
			}
			for (int q = t; q < M; q += TGPU_THREADS) {
				const int base = axis == 0 ? q * N : (axis == 1 ? (q % N) + (q / N) * M : q);
#pragma unroll
				for (int j = 0; j < N; j++) v[j] = W[base + j * step];
				// fast forms for DST-II / III and DST-IV / DCT-IV, the dense matrix for DCT-II / III (kernels.cuh)
				general_transform<N>(mats, pass < 3 ? ak.fwd : ak.inv, v, W + base, step, mg);
			}
			__syncthreads();
		}
		if (write_u)
			for (int i = t; i < NC; i += TGPU_THREADS) u[(size_t) p * NC + i] = W[i];
		if (emit)
			for (int i = t; i < 6 * M; i += TGPU_THREADS) {
				int c[3];
				face_cell<3, N>(i / M, i % M, c);
				Fout[(size_t) p * 6 * M + i] = W[(c[2] * N + c[1]) * N + c[0]];
			}
		__syncthreads(); // the block is refilled next
	}
}

// ---------------------------------------------------------------------------------------------
// operator apply / residual for 32^3 patches: one (patch, z slab) item per iteration.  Tile with a
// ghost layer [10][34][36] (interior rows start 16-byte aligned at column 2).  MODE 0: out = A u,
// MODE 1: out = f - A u.  Ghost = 2 gamma - a (neighbour), -a (Dirichlet), +a (Neumann), StarPatchOp.h:46-64.
// ---------------------------------------------------------------------------------------------
constexpr int    A32_ROW = 36, A32_PL = 34 * A32_ROW, A32_TILE = 10 * A32_PL;
constexpr size_t apply3d32_smem_bytes() { return sizeof(double) * A32_TILE; }
__device__ __forceinline__ int a32_idx(int x, int y, int zl) { return (zl + 1) * A32_PL + (y + 1) * A32_ROW + (x + 2); }

template <int MODE>
__global__ void __launch_bounds__(TGPU_THREADS, 2)
apply3d32_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ u, const double *__restrict__ f,
                 const double *__restrict__ F, double *__restrict__ out)
{
	constexpr int N = 32, M = N * N, NC = N * N * N, SLAB = 8 * M;
	extern __shared__ __align__(16) double U[];
	const int t = threadIdx.x, lo = t & 31, hi = t >> 5;
	pdl_launch_dependents();
	pdl_wait();
	const int nitems = (P - p0) * 4;
	for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
		const int        p  = p0 + (it >> 2), zs = it & 3;
		const PatchMeta &pm = meta[p];
		const double *   up = u + (size_t) p * NC;
		// interior of the slab and, where they exist inside the patch, the planes below and above it
#pragma unroll
		for (int i = 0; i < 16; i++) {
			const int c = t + TGPU_THREADS * i, row = c >> 4;
			cp_async16(U + a32_idx((c & 15) * 2, row & 31, row >> 5), up + zs * SLAB + c * 2);
		}
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int c = t + TGPU_THREADS * i, y = (c >> 4) & 31, which = c >> 9; // 512 chunks per plane, two planes
			const int z = which == 0 ? zs * 8 - 1 : zs * 8 + 8;
			if (z >= 0 && z < N) cp_async16(U + a32_idx((c & 15) * 2, y, which == 0 ? -1 : 8), up + (size_t) z * M + y * N + (c & 15) * 2);
		}
		cp_async_commit();
		const FaceVals<3, N, FV_PLAIN> fv{F, nullptr, meta};
		auto ghost = [&](int s, int m) {
			const double a = fv.get(p, 0, -1, s, m);
			if (pm.nbr_type[s] == NBR_NONE) return ((pm.neumann >> s) & 1) ? a : -a;
			return 2.0 * gamma_entry(pm, p, s, m, fv) - a;
		};
		{
			const int zl = hi, m = lo + N * (zs * 8 + zl);
			U[a32_idx(-1, lo, zl)] = ghost(0, m); // x faces: entry (y, z)
			U[a32_idx(N, lo, zl)]  = ghost(1, m);
			U[a32_idx(lo, -1, zl)] = ghost(2, m); // y faces: entry (x, z)
			U[a32_idx(lo, N, zl)]  = ghost(3, m);
			if (zs == 0 || zs == 3) {
				const int s = zs == 0 ? 4 : 5, gz = zs == 0 ? -1 : 8;
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const int mm                     = t + TGPU_THREADS * i;
					U[a32_idx(mm & 31, mm >> 5, gz)] = ghost(s, mm);
				}
			}
		}
		cp_async_wait<0>();
		__syncthreads();
		const double inv_h2 = pm.inv_h2;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int x = lo, y = hi + 8 * j;
			double    lo_v = U[a32_idx(x, y, -1)], ce = U[a32_idx(x, y, 0)];
#pragma unroll
			for (int k = 0; k < 8; k++) {
				const double hi_v = U[a32_idx(x, y, k + 1)];
				const double acc  = (U[a32_idx(x - 1, y, k)] - 2 * ce + U[a32_idx(x + 1, y, k)])
				                   + (U[a32_idx(x, y - 1, k)] - 2 * ce + U[a32_idx(x, y + 1, k)]) + (lo_v - 2 * ce + hi_v);
				const size_t o = (size_t) p * NC + (size_t) (zs * 8 + k) * M + y * N + x;
				out[o]         = (MODE == 0) ? acc * inv_h2 : __ldg(f + o) - acc * inv_h2;
				lo_v           = ce;
				ce             = hi_v;
			}
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------
// residual + restriction from face data for patches whose face (M entries) is larger than the block
// (see face_residual_restrict_kernel for the identity used); one patch per CTA, R[6][M] in dynamic smem
// ---------------------------------------------------------------------------------------------
// one patch, general form (any neighbour types, patches present on both levels); R = [S][M] staging in shared memory;
// every thread of the CTA calls; ends with a barrier
template <int D, int N, bool DIFF>
__device__ __forceinline__ void frr_big_patch(const PatchMeta *__restrict__ meta, int p, double *R, const double *__restrict__ Fnew,
                                              const double *__restrict__ Fold, double *__restrict__ coarse)
{
	using G         = Geo<D, N>;
	constexpr int H = N / 2, M = G::M;
	const int        t      = threadIdx.x;
	const PatchMeta &pm     = meta[p];
	const int        orth   = pm.orth_on_parent;
	const double     cfac   = 2.0 * pm.inv_h2;
	double *         dst    = coarse + (size_t) pm.parent_idx * G::NC;
	const FaceVals<D, N, DIFF ? FV_DIFF : FV_NEG> fv{Fnew, Fold, meta};
	for (int i = t; i < G::S * M; i += TGPU_THREADS) {
		const int s = i / M, m = i % M;
		R[i]        = cfac * gamma_entry(pm, p, s, m, fv);
	}
	__syncthreads();
	auto Rp = [&](int s, int m) { return R[s * M + m]; };
	if (orth < 0) { // patch present on both levels: coarse = r (dense, zero in the interior)
		for (int c = t; c < G::NC; c += TGPU_THREADS) {
			const int x = c % N, y = (c / N) % N, k = (D == 2) ? 0 : c / (N * N);
			double    v = 0.0;
			if (D == 2) {
				if (x == 0) v += Rp(0, y);
				if (x == N - 1) v += Rp(1, y);
				if (y == 0) v += Rp(2, x);
				if (y == N - 1) v += Rp(3, x);
			} else {
				if (x == 0) v += Rp(0, y + N * k);
				if (x == N - 1) v += Rp(1, y + N * k);
				if (y == 0) v += Rp(2, x + N * k);
				if (y == N - 1) v += Rp(3, x + N * k);
				if (k == 0) v += Rp(4, x + N * y);
				if (k == N - 1) v += Rp(5, x + N * y);
			}
			dst[c] = v;
		}
	} else {
		const int     ox = (orth & 1) * H, oy = ((orth >> 1) & 1) * H, oz = (D == 2) ? 0 : ((orth >> 2) & 1) * H;
		constexpr int CC = G::NC >> D;
		for (int c = t; c < CC; c += TGPU_THREADS) {
			const int X = c % H, Y = (c / H) % H, Z = (D == 2) ? 0 : c / (H * H);
			double    v = 0.0;
			if (D == 2) {
				auto blk = [&](int s, int I) { return (Rp(s, 2 * I) + Rp(s, 2 * I + 1)) / 4.0; };
				if (X == 0) v += blk(0, Y);
				if (X == H - 1) v += blk(1, Y);
				if (Y == 0) v += blk(2, X);
				if (Y == H - 1) v += blk(3, X);
				dst[(Y + oy) * N + (X + ox)] = v;
			} else {
				auto blk = [&](int s, int I, int J) {
					const int b = 2 * I + N * 2 * J;
					return ((Rp(s, b) + Rp(s, b + 1)) + (Rp(s, b + N) + Rp(s, b + N + 1))) / 8.0;
				};
				if (X == 0) v += blk(0, Y, Z);
				if (X == H - 1) v += blk(1, Y, Z);
				if (Y == 0) v += blk(2, X, Z);
				if (Y == H - 1) v += blk(3, X, Z);
				if (Z == 0) v += blk(4, X, Y);
				if (Z == H - 1) v += blk(5, X, Y);
				dst[((Z + oz) * N + (Y + oy)) * N + (X + ox)] = v;
			}
		}
	}
	__syncthreads();
}
template <int D, int N, bool DIFF, bool HALO = false>
__global__ void __launch_bounds__(TGPU_THREADS)
face_residual_restrict_big_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ Fnew,
                                  const double *__restrict__ Fold, double *__restrict__ coarse, HaloSync hs = HaloSync{})
{
	using G = Geo<D, N>;
	extern __shared__ __align__(16) double R[]; // [S][M]
	const int t = threadIdx.x;
	pdl_launch_dependents();
	pdl_wait();
	if (HALO) halo_push<D, N>(hs, meta);
	bool halo_ok = false;
	for (int g = blockIdx.x; g < P - p0; g += gridDim.x) {
		const int p = p0 + g;
		if (HALO) halo_wait_cta(hs, p, halo_ok);
		if (D == 3 && N == 32) {
			// Refined patches whose six sides have same-level neighbours (or none) - every patch of a uniform level - take a
			// leaner path, like face_residual_restrict16_kernel: one thread per COARSE face entry (6 x 256 per patch) loads
			// its 2 x 2 block of the patch's and the neighbour's slices with four 128-bit loads and stores the block
			// average; the 16^3 coarse cells under the patch are then assembled two per thread and stored as double2.
			// Same expressions and summation order as frr_big_patch.
			const PatchMeta &pm   = meta[p];
			bool             fast = pm.orth_on_parent >= 0;
#pragma unroll
			for (int s = 0; s < 6; s++) fast = fast && pm.nbr_type[s] <= NBR_NORMAL;
			if (fast) { // CTA-uniform
				const double cfac = 2.0 * pm.inv_h2;
				double *     Rc   = R; // [6][256]
#pragma unroll
				for (int k = 0; k < 6; k++) {
					const int e = t + TGPU_THREADS * k, s = k, c = t, m0 = 2 * (c & 15) + 64 * (c >> 4);
					double    val = 0.0;
					if (pm.nbr_type[s] == NBR_NORMAL) {
						const size_t oa = ((size_t) p * 6 + s) * 1024 + m0, ob = ((size_t) pm.nbr_idx[s][0] * 6 + (s ^ 1)) * 1024 + m0;
						auto ld = [&](size_t o) {
							double2 v = __ldg(reinterpret_cast<const double2 *>(Fnew + o));
							if (DIFF) {
								const double2 w = __ldg(reinterpret_cast<const double2 *>(Fold + o));
								v.x = w.x - v.x, v.y = w.y - v.y;
							} else {
								v.x = -v.x, v.y = -v.y;
							}
							return v;
						};
						const double2 a0 = ld(oa), a1 = ld(oa + 32), b0 = ld(ob), b1 = ld(ob + 32);
						const double  r00 = cfac * (0.5 * a0.x + 0.5 * b0.x), r01 = cfac * (0.5 * a0.y + 0.5 * b0.y);
						const double  r10 = cfac * (0.5 * a1.x + 0.5 * b1.x), r11 = cfac * (0.5 * a1.y + 0.5 * b1.y);
						val               = ((r00 + r01) + (r10 + r11)) / 8.0;
					}
					Rc[e] = val;
				}
				__syncthreads();
				const int orth = pm.orth_on_parent;
				const int ox = (orth & 1) * 16, oy = ((orth >> 1) & 1) * 16, oz = ((orth >> 2) & 1) * 16;
				double *  dst = coarse + (size_t) pm.parent_idx * G::NC;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const int q = t + TGPU_THREADS * k;                          // pair index: cells (X, Y, Z), (X + 1, Y, Z)
					const int X = (q & 7) * 2, Y = (q >> 3) & 15, Z = q >> 7;
					double    v[2];
#pragma unroll
					for (int i = 0; i < 2; i++) {
						const int x = X + i;
						double    a = 0.0;
						if (x == 0) a += Rc[0 * 256 + Y + 16 * Z];
						if (x == 15) a += Rc[1 * 256 + Y + 16 * Z];
						if (Y == 0) a += Rc[2 * 256 + x + 16 * Z];
						if (Y == 15) a += Rc[3 * 256 + x + 16 * Z];
						if (Z == 0) a += Rc[4 * 256 + x + 16 * Y];
						if (Z == 15) a += Rc[5 * 256 + x + 16 * Y];
						v[i] = a;
					}
					*reinterpret_cast<double2 *>(dst + ((Z + oz) * 32 + (Y + oy)) * 32 + (X + ox)) = make_double2(v[0], v[1]);
				}
				__syncthreads();
				continue;
			}
		}
		frr_big_patch<D, N, DIFF>(meta, p, R, Fnew, Fold, coarse);
	}
	if (HALO) halo_finish(hs);
}
} // namespace tgpu
