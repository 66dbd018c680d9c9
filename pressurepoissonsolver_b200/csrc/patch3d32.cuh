// patch3d32.cuh - kernels for D = 3, N = 32 patches (BASELINE config D).  Included by kernels.cuh.
//
// A 32^3 patch is 256 KB of fp64: more than one SM's shared memory, so the one-CTA-one-tile scheme of the
// smaller sizes does not apply.  Here one CTA still owns a whole patch but walks it in four slabs of
// eight z planes (64 KB tile):
//   smooth3d32_kernel   phase A  per z slab: f -> tile, subtract (2/h^2) gamma on the boundary cells,
//                                DST-II along x and y, result -> a CTA-private 256 KB scratch block
//                       phase B  DST-II along z, eigenvalue division, DST-III along z: register-only,
//                                pencils read/written in place in the scratch block (coalesced)
//                       phase C  per z slab: DST-III along y and x, u and its boundary slices -> global
//                       The scratch block is reused for every patch the persistent CTA processes, so it
//                       lives in L2 (296 CTAs x 256 KB = 78 MB of the 126 MB): HBM sees f once and u once,
//                       exactly like the single-tile kernels.
//   apply3d32_kernel    operator / residual with fused ghost fill, one (patch, z slab) item at a time
//   face_residual_restrict32_kernel   residual + restriction from face data (see kernels.cuh)
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
template <int D, int N, bool DIFF>
__global__ void __launch_bounds__(TGPU_THREADS)
face_residual_restrict_big_kernel(const PatchMeta *__restrict__ meta, int p0, int P, const double *__restrict__ Fnew,
                                  const double *__restrict__ Fold, double *__restrict__ coarse)
{
	using G         = Geo<D, N>;
	constexpr int H = N / 2, M = G::M;
	extern __shared__ __align__(16) double R[]; // [S][M]
	const int t = threadIdx.x;
	pdl_launch_dependents();
	pdl_wait();
	for (int g = blockIdx.x; g < P - p0; g += gridDim.x) {
		const int        p      = p0 + g;
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
}
} // namespace tgpu
