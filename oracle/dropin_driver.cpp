/* oracle/dropin_driver.cpp - the drop-in build: the adaptors of include/tgpu_thunderegg.hpp compiled against the
 * reference's REAL headers (where they lie under /root/reference, see oracle/Makefile) and driven by the reference's
 * OWN code: Tree / ThundereggDomGen build the domains, Init::initDirichlet{,2d} fills the right-hand side
 * (apps/shared/Init.cpp:152-245,305-361), GMG::VCycle / GMG::WCycle (GMG/VCycle.h, GMG/WCycle.h over GMG/Cycle.h) and
 * BiCGStab<D>::solve (BiCGStab.h:45-106) run unmodified - every smoother / operator / transfer / vector op they call is
 * a B200 kernel behind the C ABI.
 *
 * TEST INFRASTRUCTURE (tests/test_dropin.py compares its outputs with the golden vectors); needs a GPU to run.
 *
 * usage: dropin_gmg D mesh.bin divide n OPTS outdir      OPTS: "-" or the "@" options of ref_gmg without the "@"
 * writes into outdir: rhs_f.bin, rhs_exact.bin, vcycle_plugin.bin (reference cycle object over the adaptors),
 *   vcycle_fused.bin (TgpuCycle: one tgpu_vcycle call), bicgstab_plugin.bin, bicgstab_fused.bin; prints one JSON line
 */
#include <array>
#include <bitset>
#include <cmath>
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

#define private public
#define protected public
#include <Thunderegg/ThundereggDomGen.h>
#undef private
#undef protected
#include <Init.h>
#include <tgpu_thunderegg.hpp>

using namespace std;

template <size_t D> struct Problem;
template <> struct Problem<3> {
	static void init(Domain<3> &d, Vec f, Vec e)
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
};
template <> struct Problem<2> {
	static void init(Domain<2> &d, Vec f, Vec e)
	{
		auto ffun = [](double x, double y) { return (double) (-5 * M_PI * M_PI * sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		auto gfun = [](double x, double y) { return (double) (sinl(M_PI * y) * cosl(2 * M_PI * x)); };
		Init::initDirichlet2d(d, f, e, ffun, gfun);
	}
};

template <size_t D> static void dump(const string &path, const shared_ptr<Vector<D>> &v, int npatch, int cells_per_patch)
{
	/* through the adaptor's getLocalData / LocalDataManager hook, patch by patch */
	ofstream out(path, ios::binary);
	for (int p = 0; p < npatch; p++) {
		const LocalData<D> ld = static_cast<const Vector<D> &>(*v).getLocalData(p);
		array<int, D>      zero;
		zero.fill(0);
		out.write((const char *) ld.getPtr(zero), (streamsize) cells_per_patch * 8);
	}
}

template <size_t D> static int run(char **argv)
{
	using namespace tgpu_te;
	const string mesh = argv[2], optstr = argv[5], outdir = argv[6];
	const int    divide = atoi(argv[3]), n = atoi(argv[4]);
	GMG::CycleOpts opts;
	if (optstr != "-") {
		stringstream ss(optstr);
		string       o;
		while (getline(ss, o, ',')) {
			if (o == "W" || o == "V") opts.cycle_type = o;
			else if (o.rfind("pre=", 0) == 0) opts.pre_sweeps = stoi(o.substr(4));
			else if (o.rfind("post=", 0) == 0) opts.post_sweeps = stoi(o.substr(5));
			else if (o.rfind("mid=", 0) == 0) opts.mid_sweeps = stoi(o.substr(4));
			else if (o.rfind("coarse=", 0) == 0) opts.coarse_sweeps = stoi(o.substr(7));
			else if (o.rfind("max_levels=", 0) == 0) opts.max_levels = stoi(o.substr(11));
			else if (o.rfind("ppp=", 0) == 0) opts.patches_per_proc = stod(o.substr(4));
			else { cerr << "unknown option " << o << "\n"; return 2; }
		}
	}
	/* the reference's own mesh pipeline (apps/3d/steady.cpp:211-215,292-310) */
	Tree<D> t(mesh);
	for (int i = 0; i < divide; i++) t.refineLeaves();
	array<int, D> ns;
	ns.fill(n);
	shared_ptr<ThundereggDomGen<D>> dcg(new ThundereggDomGen<D>(t, ns, false));
	vector<shared_ptr<Domain<D>>>   domains;
	domains.push_back(dcg->getFinestDomain());
	while (dcg->hasCoarserDomain()) domains.push_back(dcg->getCoarserDomain());
	vector<int> global_patches;
	for (auto &d : domains) global_patches.push_back(d->getNumGlobalPatches());

	shared_ptr<Handles> hd = createHierarchy<D>(domains, n);
	const int P = domains[0]->getNumLocalPatches(), npc = domains[0]->getNumCellsInPatch();

	/* right-hand side: the reference's Init on its own host vectors, then into device vectors through Vector<D>::copy's
	 * generic getLocalData loop (the LocalDataManager path of the adaptor) */
	auto f_host = domains[0]->getNewDomainVec(), e_host = domains[0]->getNewDomainVec();
	Problem<D>::init(*domains[0], f_host->vec, e_host->vec);
	shared_ptr<VectorGenerator<D>> vg(new TgpuVG<D>(hd, 0, n));
	auto f = vg->getNewVector(), exact = vg->getNewVector(), u = vg->getNewVector(), r = vg->getNewVector();
	f->copy(f_host);
	exact->copy(e_host);
	dump<D>(outdir + "/rhs_f.bin", f, P, npc);
	dump<D>(outdir + "/rhs_exact.bin", exact, P, npc);
	const double integral = domains[0]->integrate(f); /* Domain::integrate (Domain.h:258-278) through getLocalData */
	const double integral_host = domains[0]->integrate(f_host);

	shared_ptr<Operator<D>> A(new TgpuOp<D>(hd, 0));
	/* (1) the reference's cycle object, every step a virtual call into an adaptor */
	shared_ptr<GMG::Cycle<D>> plugin_cycle = getCycle<D>(hd, n, opts, global_patches);
	plugin_cycle->apply(f, u);
	dump<D>(outdir + "/vcycle_plugin.bin", u, P, npc);
	/* (2) the same cycle as one ABI call */
	shared_ptr<Operator<D>> fused_cycle(new TgpuCycle<D>(hd, opts));
	fused_cycle->apply(f, u);
	dump<D>(outdir + "/vcycle_fused.bin", u, P, npc);
	/* (3) the reference's BiCGStab with either as the right preconditioner (apps/3d/steady.cpp:522) */
	u->set(0);
	const int its_plugin = BiCGStab<D>::solve(vg, A, u, f, plugin_cycle, 100, 1e-12);
	A->apply(u, r);
	r->scaleThenAdd(-1, f);
	const double res_plugin = r->twoNorm() / f->twoNorm();
	dump<D>(outdir + "/bicgstab_plugin.bin", u, P, npc);
	u->set(0);
	const int its_fused = BiCGStab<D>::solve(vg, A, u, f, fused_cycle, 100, 1e-12);
	A->apply(u, r);
	r->scaleThenAdd(-1, f);
	const double res_fused = r->twoNorm() / f->twoNorm();
	dump<D>(outdir + "/bicgstab_fused.bin", u, P, npc);
	/* a vector that is not a device vector must be refused by the kernels the way the reference refuses foreign
	 * vectors (throw 3, SchurHelper.h:129) */
	int threw = 0;
	try {
		A->apply(f_host, u);
	} catch (int e) {
		threw = e;
	}
	printf("{\"levels\": %zu, \"patches\": %d, \"its_plugin\": %d, \"its_fused\": %d, \"res_plugin\": %.3e, \"res_fused\": %.3e, "
	       "\"integral\": %.17g, \"integral_host\": %.17g, \"foreign_vector_throw\": %d}\n",
	       domains.size(), P, its_plugin, its_fused, res_plugin, res_fused, integral, integral_host, threw);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 7) {
		cerr << "usage: dropin_gmg D mesh.bin divide n OPTS|- outdir\n";
		return 2;
	}
	PetscInitialize(nullptr, nullptr, nullptr, nullptr);
	try {
		return atoi(argv[1]) == 2 ? run<2>(argv) : run<3>(argv);
	} catch (int e) {
		cerr << "dropin_gmg: exception " << e << ": " << tgpu_last_error() << "\n";
		return 1;
	}
}
