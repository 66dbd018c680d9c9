/* Implementation of the single-rank PETSc / FFTW / BLAS stand-ins declared in this
 * directory.  TEST INFRASTRUCTURE: lets the reference's own GMG sources be compiled
 * unmodified into oracle/_ref (see oracle/Makefile).  Never linked into the product. */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fftw3.h>
#include <petscsys.h>

static void shim_unreachable(const char *what)
{
	std::fprintf(stderr, "oracle shim: %s is a link stub and must not be reached\n", what);
	std::abort();
}
PetscErrorCode PetscObjectDestroy(PetscObject *o)
{
	if (o && *o) { delete *o; *o = nullptr; }
	return 0;
}
PetscErrorCode PetscInitialize(int *, char ***, const char *, const char *) { return 0; }
PetscErrorCode PetscFinalize() { return 0; }

/* ---- Vec ---- */
PetscErrorCode VecCreateMPI(MPI_Comm, PetscInt n, PetscInt, Vec *v)
{
	*v = new _p_Vec();
	(*v)->data.assign((size_t) n, 0.0);
	return 0;
}
PetscErrorCode VecCreateSeq(MPI_Comm c, PetscInt n, Vec *v) { return VecCreateMPI(c, n, n, v); }
PetscErrorCode VecDestroy(Vec *v)
{
	if (v && *v) { delete *v; *v = nullptr; }
	return 0;
}
PetscErrorCode VecGetArray(Vec v, double **a) { *a = v->data.data(); return 0; }
PetscErrorCode VecGetArrayRead(Vec v, const double **a) { *a = v->data.data(); return 0; }
PetscErrorCode VecRestoreArray(Vec, double **) { return 0; }
PetscErrorCode VecRestoreArrayRead(Vec, const double **) { return 0; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = (PetscInt) v->data.size(); return 0; }
PetscErrorCode VecSet(Vec v, double a) { for (double &x : v->data) x = a; return 0; }
PetscErrorCode VecScale(Vec v, double a) { for (double &x : v->data) x *= a; return 0; }
PetscErrorCode VecShift(Vec v, double a) { for (double &x : v->data) x += a; return 0; }
PetscErrorCode VecAXPY(Vec y, double a, Vec x)
{
	for (size_t i = 0; i < y->data.size(); i++) y->data[i] += a * x->data[i];
	return 0;
}
PetscErrorCode VecAYPX(Vec y, double a, Vec x)
{
	for (size_t i = 0; i < y->data.size(); i++) y->data[i] = a * y->data[i] + x->data[i];
	return 0;
}

/* ---- IS / VecScatter / AO ---- */
PetscErrorCode ISCreateBlock(MPI_Comm, PetscInt bs, PetscInt n, const PetscInt idx[], PetscCopyMode, IS *is)
{
	*is       = new _p_IS();
	(*is)->bs = bs;
	(*is)->idx.assign(idx, idx + n);
	return 0;
}
PetscErrorCode VecScatterCreate(Vec, IS ix, Vec, IS iy, VecScatter *sc)
{
	if (iy != nullptr) shim_unreachable("VecScatterCreate with a destination IS");
	*sc        = new _p_VecScatter();
	(*sc)->bs  = ix->bs;
	(*sc)->idx = ix->idx;
	return 0;
}
PetscErrorCode VecScatterBegin(VecScatter sc, Vec from, Vec to, InsertMode im, ScatterMode sm)
{
	const int bs = sc->bs;
	if (sm == SCATTER_FORWARD) {
		/* from = x (global), to = y (dist) */
		for (size_t j = 0; j < sc->idx.size(); j++) {
			const double *src = &from->data[(size_t) sc->idx[j] * bs];
			double *      dst = &to->data[j * bs];
			if (im == ADD_VALUES) for (int k = 0; k < bs; k++) dst[k] += src[k];
			else                  for (int k = 0; k < bs; k++) dst[k] = src[k];
		}
	} else {
		/* from = y (dist), to = x (global) */
		for (size_t j = 0; j < sc->idx.size(); j++) {
			const double *src = &from->data[j * bs];
			double *      dst = &to->data[(size_t) sc->idx[j] * bs];
			if (im == ADD_VALUES) for (int k = 0; k < bs; k++) dst[k] += src[k];
			else                  for (int k = 0; k < bs; k++) dst[k] = src[k];
		}
	}
	return 0;
}
PetscErrorCode VecScatterEnd(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
PetscErrorCode AOCreateMapping(MPI_Comm, PetscInt n, const PetscInt app[], const PetscInt petsc[], AO *ao)
{
	*ao = new _p_AO();
	for (int i = 0; i < n; i++) (*ao)->map[app[i]] = petsc[i];
	return 0;
}
PetscErrorCode AOApplicationToPetsc(AO ao, PetscInt n, PetscInt ia[])
{
	for (int i = 0; i < n; i++) {
		auto it = ao->map.find(ia[i]);
		ia[i]   = it == ao->map.end() ? -1 : it->second;
	}
	return 0;
}

/* ---- Mat: rows of (column -> value) maps, one rank ---- */
PetscErrorCode MatCreate(MPI_Comm, Mat *m) { *m = new _p_Mat(); return 0; }
PetscErrorCode MatSetSizes(Mat m, PetscInt local_rows, PetscInt, PetscInt, PetscInt) { m->rows.assign(local_rows, std::map<int, double>()); return 0; }
PetscErrorCode MatSetType(Mat, MatType) { return 0; }
PetscErrorCode MatMPIAIJSetPreallocation(Mat, PetscInt, const PetscInt *, PetscInt, const PetscInt *) { return 0; }
PetscErrorCode MatSetValues(Mat m, PetscInt nr, const PetscInt *r, PetscInt nc, const PetscInt *c, const PetscScalar *v, InsertMode mode)
{
	for (PetscInt i = 0; i < nr; i++) {
		if (r[i] < 0 || r[i] >= (PetscInt) m->rows.size()) shim_unreachable("MatSetValues: row outside the local range");
		for (PetscInt j = 0; j < nc; j++) {
			if (mode == ADD_VALUES) m->rows[r[i]][c[j]] += v[i * nc + j];
			else m->rows[r[i]][c[j]] = v[i * nc + j];
		}
	}
	return 0;
}
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatMult(Mat m, Vec x, Vec y)
{
	if (x->data.size() != m->rows.size() || y->data.size() != m->rows.size()) shim_unreachable("MatMult: size mismatch");
	for (size_t i = 0; i < m->rows.size(); i++) {
		double s = 0;
		for (const auto &e : m->rows[i]) {
			if (e.first < 0 || e.first >= (int) x->data.size()) shim_unreachable("MatMult: column outside the vector");
			s += e.second * x->data[e.first];
		}
		y->data[i] = s;
	}
	return 0;
}
PetscErrorCode MatGetRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **) { shim_unreachable("MatGetRow"); return 1; }
PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **) { shim_unreachable("MatRestoreRow"); return 1; }
PetscErrorCode PCSetType(PC, PCType) { shim_unreachable("PCSetType"); return 1; }
PetscErrorCode PCShellSetContext(PC, void *) { shim_unreachable("PCShellSetContext"); return 1; }
PetscErrorCode PCShellGetContext(PC, void **) { shim_unreachable("PCShellGetContext"); return 1; }
PetscErrorCode PCShellSetApply(PC, PetscErrorCode (*)(PC, Vec, Vec)) { shim_unreachable("PCShellSetApply"); return 1; }

/* ---- BLAS dgemv (only the call shape DftPatchSolver issues: trans='T', alpha=1, beta=0) ---- */
extern "C" void dgemv_(char &trans, int &m, int &n, double &alpha, double *a, int &lda, double *x,
                       int &incx, double &beta, double *y, int &incy)
{
	if (trans != 'T' && trans != 't') shim_unreachable("dgemv_ with trans != 'T'");
	/* y_i = beta*y_i + alpha * sum_j A(j,i) x_j,  A column-major m x n */
	for (int i = 0; i < n; i++) {
		const double *col = a + (size_t) i * lda;
		double        acc = 0;
		for (int j = 0; j < m; j++) acc += col[j] * x[(size_t) j * incx];
		double *yi = y + (size_t) i * incy;
		*yi        = (beta == 0 ? 0 : beta * *yi) + alpha * acc;
	}
}

/* ---- FFTW r2r (naive separable evaluation of the manual's definitions) ---- */
struct fftw_plan_s {
	int                        rank;
	std::vector<int>           n;
	double *                   in;
	double *                   out;
	std::vector<fftw_r2r_kind> kind;
	std::vector<std::vector<double>> mats; /* mats[d][k*n+j]: Y_k = sum_j mats[k*n+j] X_j */
};
static std::vector<double> r2r_matrix(fftw_r2r_kind kind, int n)
{
	std::vector<double> m((size_t) n * n, 0.0);
	for (int k = 0; k < n; k++) {
		for (int j = 0; j < n; j++) {
			double v = 0;
			switch (kind) {
				case FFTW_REDFT10: v = 2 * std::cos(M_PI * (j + 0.5) * k / n); break;
				case FFTW_REDFT01: v = j == 0 ? 1.0 : 2 * std::cos(M_PI * j * (k + 0.5) / n); break;
				case FFTW_REDFT11: v = 2 * std::cos(M_PI * (j + 0.5) * (k + 0.5) / n); break;
				case FFTW_RODFT10: v = 2 * std::sin(M_PI * (j + 0.5) * (k + 1) / n); break;
				case FFTW_RODFT01:
					v = j == n - 1 ? ((k % 2) ? -1.0 : 1.0) : 2 * std::sin(M_PI * (j + 1) * (k + 0.5) / n);
					break;
				case FFTW_RODFT11: v = 2 * std::sin(M_PI * (j + 0.5) * (k + 0.5) / n); break;
				default: shim_unreachable("fftw r2r kind outside the six the reference plans");
			}
			m[(size_t) k * n + j] = v;
		}
	}
	return m;
}
fftw_plan fftw_plan_r2r(int rank, const int *n, double *in, double *out, const fftw_r2r_kind *kind, unsigned)
{
	fftw_plan p = new fftw_plan_s();
	p->rank     = rank;
	p->n.assign(n, n + rank);
	p->in  = in;
	p->out = out;
	p->kind.assign(kind, kind + rank);
	for (int d = 0; d < rank; d++) p->mats.push_back(r2r_matrix(kind[d], n[d]));
	return p;
}
void fftw_execute(const fftw_plan p)
{
	size_t total = 1;
	for (int d = 0; d < p->rank; d++) total *= (size_t) p->n[d];
	std::vector<double> a(p->in, p->in + total), b(total);
	/* row-major: dimension rank-1 is contiguous */
	size_t stride = 1;
	for (int d = p->rank - 1; d >= 0; d--) {
		const int                  nd = p->n[d];
		const std::vector<double> &m  = p->mats[d];
		for (size_t base = 0; base < total; base++) {
			if ((base / stride) % nd != 0) continue;
			for (int k = 0; k < nd; k++) {
				double acc = 0;
				for (int j = 0; j < nd; j++) acc += m[(size_t) k * nd + j] * a[base + j * stride];
				b[base + k * stride] = acc;
			}
		}
		a.swap(b);
		stride *= nd;
	}
	std::memcpy(p->out, a.data(), total * sizeof(double));
}
void fftw_destroy_plan(fftw_plan p) { delete p; }
