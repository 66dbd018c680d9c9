/* oracle/ref_driver.cpp - command-line driver around the UNMODIFIED reference GMG sources
 * (compiled where they lie under /root/reference, see oracle/Makefile) so that its operator,
 * smoother, transfer operators, V-cycle and BiCGStab can be run on raw input files.
 *
 * TEST INFRASTRUCTURE: this program is the parity oracle ("oracle/_ref/ref_gmg") and the CPU
 * baseline binary.  It is never linked into, or called by, the product path.
 *
 * The set-up mirrors apps/3d/steady.cpp:211-215,292-310,482 and apps/2d/steady.cpp (Tree load,
 * refineLeaves x divide, ThundereggDomGen, patch solver, GMG::CycleFactory{2,3}d::getCycle) and
 * the RHS follows apps/3d/steady.cpp:253-265 / apps/2d/steady.cpp:314-316 through
 * Init::initDirichlet{,2d}.
 *
 * usage: ref_gmg D mesh.bin divide n dft|fftw[-neumann][@opt,opt...] cmd [cmd...]
 *   (suffix -neumann: ThundereggDomGen(tree, ns, neumann = true), apps/3d/steady.cpp:301)
 *   (@ options fill GMG::CycleOpts, GMG/CycleOpts.h:51-80, and the patch solver's shift: W | V | pre=K | post=K | mid=K |
 *    coarse=K | max_levels=K | ppp=X (patches_per_proc) | lambda=X (FftwPatchSolver / DftPatchSolver(domain, lambda)))
 *   cmd rhsn:trig|gauss:f.bin:exact.bin  the reference's Init::initNeumann / initNeumann2d on the app's manufactured problems
 *   meta:OUT                         hierarchy metadata (format: see dump_meta)
 *   rhs:F_OUT:EXACT_OUT              trig manufactured problem, Dirichlet data folded into f
 *   apply:L:U_IN:OUT                 OUT = A_L U                (level 0 = finest)
 *   matapply:L:U_IN:OUT              3D only: OUT = (MatrixHelper(domain_L).formCRSMatrix()) U - the reference's independent ASSEMBLED
 *                                    form of the same operator (StencilHelper.h: per-side stencils incl. coarse/fine weights)
 *   smooth:L:F_IN:U_IN:OUT           one block-Jacobi sweep on level L
 *   restrict:L:FINE_IN:OUT           AvgRstr from level L to level L+1
 *   interp:L:COARSE_IN:FINE_IN:OUT   DrctIntp from level L+1 into level L (adds)
 *   vcycle:F_IN:OUT                  OUT = Cycle::apply(F)
 *   vhist:F_IN:NCYC:U_OUT:HIST_OUT   u += V(f - A u), NCYC times from u = 0; HIST = ||f-Au||_2/||f||_2
 *   bicgstab:F_IN:TOL:MAXIT:U_OUT:INFO_OUT   BiCGStab<D>::solve with Mr = V-cycle; INFO = its, relres
 *   time:REPS[:WARMUP]               times Cycle::apply on the trig RHS, prints one JSON line
 */
#include <algorithm>
#include <array>
#include <bitset>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <valarray>
#include <vector>

/* the reference keeps the level list and the finest Level private; open them for inspection */
#define private public
#define protected public
#include <Thunderegg/GMG/Cycle.h>
#include <Thunderegg/ThundereggDomGen.h>
#undef private
#undef protected
#include <Thunderegg/BiCGStab.h>
#include <Thunderegg/BilinearInterpolator.h>
#include <Thunderegg/GMG/CycleFactory2d.h>
#include <Thunderegg/GMG/CycleFactory3d.h>
#include <Thunderegg/MatrixHelper.h>
#include <Thunderegg/Operators/DomainWrapOp.h>
#include <Thunderegg/PatchSolvers/DftPatchSolver.h>
#include <Thunderegg/PatchSolvers/FftwPatchSolver.h>
#include <Thunderegg/StarPatchOp.h>
#include <Thunderegg/TriLinInterp.h>
#include <Init.h>

using namespace std;

static vector<string> split(const string &s, char c)
{
	vector<string> out;
	string         item;
	stringstream   ss(s);
	while (getline(ss, item, c)) out.push_back(item);
	return out;
}
static void read_into(const string &path, Vec v)
{
	double *p;
	int     n;
	VecGetArray(v, &p);
	VecGetLocalSize(v, &n);
	ifstream in(path, ios::binary);
	if (!in) { cerr << "cannot open " << path << "\n"; exit(2); }
	in.read((char *) p, (streamsize) n * 8);
	if (in.gcount() != (streamsize) n * 8) { cerr << "short read " << path << " want " << n << " doubles\n"; exit(2); }
}
static void write_from(const string &path, Vec v)
{
	double *p;
	int     n;
	VecGetArray(v, &p);
	VecGetLocalSize(v, &n);
	ofstream out(path, ios::binary);
	out.write((const char *) p, (streamsize) n * 8);
}

template <size_t D> struct Traits;
template <> struct Traits<3> {
	using Factory = GMG::CycleFactory3d;
	using Interp  = TriLinInterp;
	static void matapply(shared_ptr<Domain<3>> d, Vec u, Vec out)
	{
		MatrixHelper  mh(d);
		PW<Mat>       A = mh.formCRSMatrix();
		MatMult(A, u, out);
	}
	static void rhs(Domain<3> &d, Vec f, Vec e)
	{
		auto ffun = [](double x, double y, double z) {
			x += .3; y += .3; z += .3;
			return -77.0 / 36 * M_PI * M_PI * sin(M_PI * x) * cos(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
		};
		auto gfun = [](double x, double y, double z) {
			x += .3; y += .3; z += .3;
			return sin(M_PI * x) * cos(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
		};
		Init::initDirichlet(d, f, e, ffun, gfun);
	}
	/* Init::initNeumann with the manufactured problems of apps/3d/steady.cpp:230-282 ("gauss", default trig) */
	static void rhs_neumann(Domain<3> &d, Vec f, Vec e, const string &problem)
	{
		function<double(double, double, double)> ffun, gfun, nfunx, nfuny, nfunz;
		if (problem == "gauss") {
			gfun = [](double x, double y, double z) { return exp(cos(10 * M_PI * x)) - exp(cos(11 * M_PI * y)) + exp(cos(12 * M_PI * z)); };
			ffun = [](double x, double y, double z) {
				return -M_PI * M_PI
				       * (100 * exp(cos(10 * M_PI * x)) * cos(10 * M_PI * x) - 100 * exp(cos(10 * M_PI * x)) * pow(sin(10 * M_PI * x), 2)
				          - 121 * exp(cos(11 * M_PI * y)) * cos(11 * M_PI * y) + 121 * exp(cos(11 * M_PI * y)) * pow(sin(11 * M_PI * y), 2)
				          + 144 * exp(cos(12 * M_PI * z)) * cos(12 * M_PI * z) - 144 * exp(cos(12 * M_PI * z)) * pow(sin(12 * M_PI * z), 2));
			};
			nfunx = [](double x, double y, double z) { return -10 * M_PI * sin(10 * M_PI * x) * exp(cos(10 * M_PI * x)); };
			nfuny = [](double x, double y, double z) { return 11 * M_PI * sin(11 * M_PI * y) * exp(cos(11 * M_PI * y)); };
			nfunz = [](double x, double y, double z) { return -12 * M_PI * sin(12 * M_PI * z) * exp(cos(12 * M_PI * z)); };
		} else {
			ffun = [](double x, double y, double z) {
				x += .3; y += .3; z += .3;
				return -77.0 / 36 * M_PI * M_PI * sin(M_PI * x) * cos(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
			};
			gfun = [](double x, double y, double z) {
				x += .3; y += .3; z += .3;
				return sin(M_PI * x) * cos(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
			};
			nfunx = [](double x, double y, double z) {
				x += .3; y += .3; z += .3;
				return M_PI * cos(M_PI * x) * cos(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
			};
			nfuny = [](double x, double y, double z) {
				x += .3; y += .3; z += .3;
				return -2.0 / 3 * M_PI * sin(M_PI * x) * sin(2.0 / 3 * M_PI * y) * sin(5.0 / 6 * M_PI * z);
			};
			nfunz = [](double x, double y, double z) {
				x += .3; y += .3; z += .3;
				return 5.0 / 6 * M_PI * sin(M_PI * x) * cos(2.0 / 3 * M_PI * y) * cos(5.0 / 6 * M_PI * z);
			};
		}
		Init::initNeumann(d, f, e, ffun, gfun, nfunx, nfuny, nfunz);
	}
};
template <> struct Traits<2> {
	using Factory = GMG::CycleFactory2d;
	using Interp  = BilinearInterpolator;
	static void matapply(shared_ptr<Domain<2>>, Vec, Vec)
	{
		/* StencilHelper2d.h assembles a different (quadratic) coarse/fine stencil: not the matrix-free operator */
		cerr << "matapply: 3D only\n";
		exit(2);
	}
	static void rhs(Domain<2> &d, Vec f, Vec e)
	{
		auto ffun = [](double x, double y) { return (double) (-5 * M_PI * M_PI * sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		auto gfun = [](double x, double y) { return (double) (sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		Init::initDirichlet2d(d, f, e, ffun, gfun);
	}
	/* Init::initNeumann2d with the trig problem of apps/2d/steady.cpp:314-318 */
	static void rhs_neumann(Domain<2> &d, Vec f, Vec e, const string &problem)
	{
		if (problem != "trig") { cerr << "rhsn: 2D has the trig problem only\n"; exit(2); }
		auto ffun  = [](double x, double y) { return (double) (-5 * M_PI * M_PI * sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		auto gfun  = [](double x, double y) { return (double) (sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		auto nfun  = [](double x, double y) { return (double) (-2 * M_PI * sinl(M_PI * y) * sinl(2 * M_PI * x)); };
		auto nfuny = [](double x, double y) { return (double) (M_PI * cosl(M_PI * y) * cosl(2 * M_PI * x)); };
		Init::initNeumann2d(d, f, e, ffun, gfun, nfun, nfuny);
	}
};

template <size_t D> struct Ctx {
	shared_ptr<ThundereggDomGen<D>>  dcg;
	shared_ptr<GMG::Cycle<D>>        cycle;
	vector<shared_ptr<Domain<D>>>    domains; /* finest first */
	vector<shared_ptr<GMG::Level<D>>> levels;
	shared_ptr<Operator<D>>          A; /* DomainWrapOp on the finest level, as apps/3d/steady.cpp:453 */
};

/* metadata record, all little-endian:
 *  header int32: magic 0x474d4731, D, n, nlevels
 *  per level (finest first): int32 npatch; int32 ints[npatch][6 + 2D*(2 + 2*2^(D-1))];
 *                            double reals[npatch][2D]  (starts[D], spacings[D])
 *  ints per patch: id, refine_level, parent_id, orth_on_parent, parent_local_index (index of
 *  parent_id in the next coarser level's local order, -1 on the coarsest), neumann bits, then
 *  per side: type (-1 none, 0 normal, 1 coarse, 2 fine), orth_on_coarse (-1 unless coarse),
 *  ids[2^(D-1)], local_indexes[2^(D-1)] (unused slots -1). */
template <size_t D> static void dump_meta(Ctx<D> &c, int n, const string &path)
{
	ofstream      out(path, ios::binary);
	const int     nq     = 1 << (D - 1);
	const int32_t hdr[4] = {0x474d4731, (int32_t) D, n, (int32_t) c.domains.size()};
	out.write((const char *) hdr, sizeof(hdr));
	for (size_t l = 0; l < c.domains.size(); l++) {
		auto &  vec = c.domains[l]->getPatchInfoVector();
		int32_t np  = vec.size();
		out.write((const char *) &np, 4);
		vector<int32_t> ints;
		vector<double>  reals;
		for (auto &pi : vec) {
			ints.push_back(pi->id);
			ints.push_back(pi->refine_level);
			ints.push_back(pi->parent_id);
			ints.push_back(pi->orth_on_parent.toInt());
			int pl = -1;
			if (l + 1 < c.domains.size()) pl = c.domains[l + 1]->getPatchInfoMap().at(pi->parent_id)->local_index;
			ints.push_back(pl);
			ints.push_back((int32_t) pi->neumann.to_ulong());
			for (Side<D> s : Side<D>::getValues()) {
				int32_t type = -1, orth = -1;
				vector<int32_t> ids(nq, -1), loc(nq, -1);
				if (pi->hasNbr(s)) {
					switch (pi->getNbrType(s)) {
						case NbrType::Normal:
							type   = 0;
							ids[0] = pi->getNormalNbrInfo(s).id;
							loc[0] = pi->getNormalNbrInfo(s).local_index;
							break;
						case NbrType::Coarse:
							type   = 1;
							ids[0] = pi->getCoarseNbrInfo(s).id;
							loc[0] = pi->getCoarseNbrInfo(s).local_index;
							orth   = pi->getCoarseNbrInfo(s).orth_on_coarse.toInt();
							break;
						case NbrType::Fine:
							type = 2;
							for (int q = 0; q < nq; q++) {
								ids[q] = pi->getFineNbrInfo(s).ids[q];
								loc[q] = pi->getFineNbrInfo(s).local_indexes[q];
							}
							break;
					}
				}
				ints.push_back(type);
				ints.push_back(orth);
				ints.insert(ints.end(), ids.begin(), ids.end());
				ints.insert(ints.end(), loc.begin(), loc.end());
			}
			for (size_t i = 0; i < D; i++) reals.push_back(pi->starts[i]);
			for (size_t i = 0; i < D; i++) reals.push_back(pi->spacings[i]);
		}
		out.write((const char *) ints.data(), ints.size() * 4);
		out.write((const char *) reals.data(), reals.size() * 8);
	}
}

template <size_t D> static int run(int argc, char **argv)
{
	string mesh   = argv[2];
	int    divide = atoi(argv[3]);
	int    n      = atoi(argv[4]);
	string solver = argv[5];
	GMG::CycleOpts opts; /* defaults: V, 1 pre, 1 post, 1 coarse sweep, all levels */
	double         lambda = 0;
	{
		size_t at = solver.find('@');
		if (at != string::npos) {
			for (const string &o : split(solver.substr(at + 1), ',')) {
				if (o == "W" || o == "V") opts.cycle_type = o;
				else if (o.rfind("pre=", 0) == 0) opts.pre_sweeps = stoi(o.substr(4));
				else if (o.rfind("post=", 0) == 0) opts.post_sweeps = stoi(o.substr(5));
				else if (o.rfind("mid=", 0) == 0) opts.mid_sweeps = stoi(o.substr(4));
				else if (o.rfind("coarse=", 0) == 0) opts.coarse_sweeps = stoi(o.substr(7));
				else if (o.rfind("max_levels=", 0) == 0) opts.max_levels = stoi(o.substr(11));
				else if (o.rfind("ppp=", 0) == 0) opts.patches_per_proc = stod(o.substr(4));
				else if (o.rfind("lambda=", 0) == 0) lambda = stod(o.substr(7));
				else { cerr << "unknown option " << o << "\n"; return 2; }
			}
			solver = solver.substr(0, at);
		}
	}
	bool   neumann = false;
	if (solver.size() > 8 && solver.substr(solver.size() - 8) == "-neumann") {
		neumann = true;
		solver  = solver.substr(0, solver.size() - 8);
	}

	auto t0 = chrono::steady_clock::now();
	Tree<D> t(mesh);
	for (int i = 0; i < divide; i++) t.refineLeaves();
	array<int, D> ns;
	ns.fill(n);

	Ctx<D> c;
	c.dcg.reset(new ThundereggDomGen<D>(t, ns, neumann));
	shared_ptr<Domain<D>>        finest = c.dcg->getFinestDomain();
	shared_ptr<PatchOperator<D>> p_op(new StarPatchOp<D>());
	shared_ptr<IfaceInterp<D>>   p_interp(new typename Traits<D>::Interp());
	shared_ptr<PatchSolver<D>>   p_solver;
	if (solver == "fftw") p_solver.reset(new FftwPatchSolver<D>(*finest, lambda));
	else                  p_solver.reset(new DftPatchSolver<D>(*finest, lambda));
	shared_ptr<SchurHelper<D>> sch(new SchurHelper<D>(finest, p_solver, p_op, p_interp));
	c.A.reset(new DomainWrapOp<D>(sch));
	c.cycle = Traits<D>::Factory::getCycle(opts, c.dcg, p_solver, p_op, p_interp);
	for (auto &d : c.dcg->domain_list) c.domains.push_back(d);
	{
		shared_ptr<GMG::Level<D>> l = c.cycle->finest_level;
		while (l) { c.levels.push_back(l); l = l->coarser; }
	}
	double setup_s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
	if (c.levels.size() > c.domains.size()) { cerr << "level/domain count mismatch\n"; return 3; }
	if (c.levels.size() != c.domains.size() && opts.max_levels == 0 && opts.patches_per_proc == 0) { cerr << "level/domain count mismatch\n"; return 3; }

	auto newvec = [&](int l) { return c.domains[l]->getNewDomainVec(); };

	for (int a = 6; a < argc; a++) {
		vector<string> p = split(argv[a], ':');
		const string & cmd = p[0];
		if (cmd == "meta") {
			dump_meta<D>(c, n, p[1]);
		} else if (cmd == "rhs") {
			auto f = newvec(0), e = newvec(0);
			Traits<D>::rhs(*c.domains[0], f->vec, e->vec);
			write_from(p[1], f->vec);
			write_from(p[2], e->vec);
		} else if (cmd == "rhsn") { /* rhsn:problem:f:exact  -> Init::initNeumann, then prints integrate(f) / volume (apps/3d/steady.cpp:330-334) */
			auto f = newvec(0), e = newvec(0);
			Traits<D>::rhs_neumann(*c.domains[0], f->vec, e->vec, p[1]);
			write_from(p[2], f->vec);
			write_from(p[3], e->vec);
			printf("{\"fdiff\": %.17g, \"volume\": %.17g}\n", c.domains[0]->integrate(f) / c.domains[0]->volume(), c.domains[0]->volume());
		} else if (cmd == "apply") {
			int  l = stoi(p[1]);
			auto u = newvec(l), o = newvec(l);
			read_into(p[2], u->vec);
			c.levels[l]->getOperator().apply(u, o);
			write_from(p[3], o->vec);
		} else if (cmd == "matapply") {
			int  l = stoi(p[1]);
			auto u = newvec(l), o = newvec(l);
			read_into(p[2], u->vec);
			Traits<D>::matapply(c.domains[l], u->vec, o->vec);
			write_from(p[3], o->vec);
		} else if (cmd == "smooth") {
			int  l = stoi(p[1]);
			auto f = newvec(l), u = newvec(l);
			read_into(p[2], f->vec);
			read_into(p[3], u->vec);
			c.levels[l]->getSmoother().smooth(f, u);
			write_from(p[4], u->vec);
		} else if (cmd == "restrict") {
			int  l = stoi(p[1]);
			auto fine = newvec(l), coarse = newvec(l + 1);
			read_into(p[2], fine->vec);
			c.levels[l]->getRestrictor().restrict(coarse, fine);
			write_from(p[3], coarse->vec);
		} else if (cmd == "interp") {
			int  l = stoi(p[1]);
			auto coarse = newvec(l + 1), fine = newvec(l);
			read_into(p[2], coarse->vec);
			read_into(p[3], fine->vec);
			c.levels[l + 1]->getInterpolator().interpolate(coarse, fine);
			write_from(p[4], fine->vec);
		} else if (cmd == "vcycle") {
			auto f = newvec(0), u = newvec(0);
			read_into(p[1], f->vec);
			c.cycle->apply(f, u);
			write_from(p[2], u->vec);
		} else if (cmd == "vhist") {
			int  ncyc = stoi(p[2]);
			auto f = newvec(0), u = newvec(0), r = newvec(0), e = newvec(0);
			read_into(p[1], f->vec);
			vector<double> hist;
			double         fn = f->twoNorm();
			for (int k = 0; k <= ncyc; k++) {
				c.A->apply(u, r);
				r->scaleThenAdd(-1, f);
				hist.push_back(r->twoNorm() / fn);
				if (k == ncyc) break;
				c.cycle->apply(r, e);
				u->add(e);
			}
			write_from(p[3], u->vec);
			ofstream out(p[4], ios::binary);
			out.write((const char *) hist.data(), hist.size() * 8);
		} else if (cmd == "bicgstab") {
			double tol   = stod(p[2]);
			int    maxit = stoi(p[3]);
			auto   f = newvec(0), u = newvec(0), r = newvec(0);
			read_into(p[1], f->vec);
			shared_ptr<VectorGenerator<D>> vg(new DomainVG<D>(c.domains[0]));
			auto   t1  = chrono::steady_clock::now();
			int    its = BiCGStab<D>::solve(vg, c.A, u, f, c.cycle, maxit, tol);
			double sec = chrono::duration<double>(chrono::steady_clock::now() - t1).count();
			c.A->apply(u, r);
			r->scaleThenAdd(-1, f);
			double info[3] = {(double) its, r->twoNorm() / f->twoNorm(), sec};
			write_from(p[4], u->vec);
			ofstream out(p[5], ios::binary);
			out.write((const char *) info, sizeof(info));
		} else if (cmd == "time") {
			int  reps = stoi(p[1]), warm = p.size() > 2 ? stoi(p[2]) : 1;
			auto f = newvec(0), e = newvec(0), u = newvec(0);
			Traits<D>::rhs(*c.domains[0], f->vec, e->vec);
			for (int k = 0; k < warm; k++) c.cycle->apply(f, u); /* warm-up */
			vector<double> secs;
			for (int k = 0; k < reps; k++) {
				auto t1 = chrono::steady_clock::now();
				c.cycle->apply(f, u);
				secs.push_back(chrono::duration<double>(chrono::steady_clock::now() - t1).count());
			}
			sort(secs.begin(), secs.end());
			double med   = secs[secs.size() / 2];
			long   cells = (long) c.domains[0]->getNumLocalPatches() * c.domains[0]->getNumCellsInPatch();
			printf("{\"cells\": %ld, \"patches\": %d, \"levels\": %zu, \"reps\": %d, \"sec_per_vcycle_median\": %.6e, "
			       "\"sec_per_vcycle_min\": %.6e, \"dof_per_s\": %.6e, \"setup_s\": %.3f, \"patch_solver\": \"%s\"}\n",
			       cells, c.domains[0]->getNumLocalPatches(), c.levels.size(), reps, med, secs[0], cells / med,
			       setup_s, solver.c_str());
		} else {
			cerr << "unknown command " << cmd << "\n";
			return 2;
		}
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 6) {
		cerr << "usage: ref_gmg D mesh.bin divide n dft|fftw cmd [cmd...]\n";
		return 2;
	}
	PetscInitialize(nullptr, nullptr, nullptr, nullptr);
	int D = atoi(argv[1]);
	return D == 2 ? run<2>(argc, argv) : run<3>(argc, argv);
}
