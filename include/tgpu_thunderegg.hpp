// tgpu_thunderegg.hpp - the drop-in binding: adaptors that DERIVE FROM THE REFERENCE'S OWN CLASSES and forward to the
// C ABI (tgpu.h).  Compile it where ThunderEgg's headers are on the include path (-I <reference>/src); nothing in this
// repository copies those headers.  With these classes the reference's unmodified drivers - GMG::VCycle / GMG::WCycle
// (GMG/VCycle.h:44-62, GMG/WCycle.h:45-68 over GMG/Cycle.h:56-126), BiCGStab<D>::solve (BiCGStab.h:45-106),
// Domain<D>::integrate (Domain.h:258-278), PetscShellCreator's thunks - run on the B200 kernels:
//
//   reference interface (src/Thunderegg/...)                 adaptor
//   Vector<D>, every op virtual          Vector.h:179-322    TgpuVector<D>      tgpu_vec_*
//     getLocalData + LocalDataManager    Vector.h:30-34,72,78-90,187-188   host mirror, acquire = D2H, release = H2D
//   VectorGenerator<D>                   Vector.h:323-327    TgpuVG<D>          tgpu_vec_create
//   Operator<D>::apply                   Operators/Operator.h:37            TgpuOp<D>          tgpu_apply
//   GMG::Smoother<D>::smooth             GMG/Smoother.h:39   TgpuSmoother<D>    tgpu_smooth  (TgpuJacobiSmoother: tgpu_smooth_jacobi)
//   GMG::Restrictor<D>::restrict         GMG/Restrictor.h:39-40             TgpuRestrictor<D>  tgpu_restrict
//   GMG::Interpolator<D>::interpolate    GMG/Interpolator.h:39-40           TgpuInterpolator<D> tgpu_prolong_add (linear: tgpu_prolong_add_linear)
//   GMG::Cycle<D> (is-an Operator<D>)    GMG/Cycle.h:34,116-126             TgpuCycle<D>       tgpu_vcycle (fused schedule, CUDA-graph replay)
//   GMG::CycleFactory3d::getCycle        GMG/CycleFactory3d.cpp:69-134      tgpu_te::getLevels / getCycle
//   Domain<D> / PatchInfo<D> metadata    Domain.h:146, PatchInfo.h:74-277   tgpu_te::flatten -> TgpuLevelDesc
//
// Error convention: the reference throws an int (`throw 3;`, SchurHelper.h:129, GMG/InterLevelComm.h:175) on type
// mismatches; the adaptors do the same for every non-zero ABI return code (message: tgpu_last_error()).
// oracle/dropin_driver.cpp builds this header against the reference's real headers and drives it with the reference's
// own Init::initDirichlet, GMG::VCycle / WCycle and BiCGStab (tests/test_dropin.py compares with the golden vectors).
#pragma once
#include <array>
#include <memory>
#include <vector>

#include <Thunderegg/BiCGStab.h>
#include <Thunderegg/Domain.h>
#include <Thunderegg/GMG/Cycle.h>
#include <Thunderegg/GMG/CycleOpts.h>
#include <Thunderegg/GMG/Interpolator.h>
#include <Thunderegg/GMG/Level.h>
#include <Thunderegg/GMG/Restrictor.h>
#include <Thunderegg/GMG/Smoother.h>
#include <Thunderegg/GMG/VCycle.h>
#include <Thunderegg/GMG/WCycle.h>
#include <Thunderegg/Operators/Operator.h>
#include <Thunderegg/Vector.h>

#include "tgpu.h"

namespace tgpu_te
{
inline void tg(int rc)
{
	if (rc != TGPU_OK) throw 3; // the reference's own convention
}

// RAII over the ABI handles, shared by every adaptor of one hierarchy
struct Handles {
	tgpu_ctx * ctx = nullptr;
	tgpu_hier *h   = nullptr;
	~Handles()
	{
		if (h) tgpu_hierarchy_destroy(h);
		if (ctx) tgpu_finalize(ctx);
	}
};

// ---------------------------------------------------------------------------------------------------------------
// Domain<D> objects of the reference (finest first) -> TgpuLevelDesc in local_index order
// ---------------------------------------------------------------------------------------------------------------
template <size_t D> struct LevelArrays {
	std::vector<double>  spacing, starts;
	std::vector<uint8_t> neumann;
	std::vector<int8_t>  nbr_type, orth_on_coarse, orth_on_parent;
	std::vector<int32_t> nbr_idx, parent_idx;
	TgpuLevelDesc        desc() const
	{
		TgpuLevelDesc d;
		d.npatch         = (int32_t) neumann.size();
		d.spacing        = spacing.data();
		d.starts         = starts.data();
		d.neumann_bits   = neumann.data();
		d.nbr_type       = nbr_type.data();
		d.nbr_idx        = nbr_idx.data();
		d.orth_on_coarse = orth_on_coarse.data();
		d.parent_idx     = parent_idx.data();
		d.orth_on_parent = orth_on_parent.data();
		return d;
	}
};
template <size_t D> LevelArrays<D> flatten(Domain<D> &dom, Domain<D> *coarser)
{
	constexpr int  Q = 1 << (D - 1);
	LevelArrays<D> a;
	for (auto &pi : dom.getPatchInfoVector()) { // local_index order, Domain.h:146
		for (size_t i = 0; i < D; i++) {
			a.spacing.push_back(pi->spacings[i]);
			a.starts.push_back(pi->starts[i]);
		}
		a.neumann.push_back((uint8_t) pi->neumann.to_ulong());
		a.orth_on_parent.push_back((int8_t) pi->orth_on_parent.toInt()); // -1: the same patch on both levels
		a.parent_idx.push_back(coarser ? coarser->getPatchInfoMap().at(pi->parent_id)->local_index : -1);
		for (Side<D> s : Side<D>::getValues()) {
			int8_t  type = TGPU_NBR_NONE, orth = -1;
			int32_t idx[Q];
			for (int q = 0; q < Q; q++) idx[q] = -1;
			if (pi->hasNbr(s)) {
				switch (pi->getNbrType(s)) {
					case NbrType::Normal:
						type   = TGPU_NBR_NORMAL;
						idx[0] = pi->getNormalNbrInfo(s).local_index;
						break;
					case NbrType::Coarse:
						type   = TGPU_NBR_COARSE;
						idx[0] = pi->getCoarseNbrInfo(s).local_index;
						orth   = (int8_t) pi->getCoarseNbrInfo(s).orth_on_coarse.toInt();
						break;
					case NbrType::Fine:
						type = TGPU_NBR_FINE;
						for (int q = 0; q < Q; q++) idx[q] = pi->getFineNbrInfo(s).local_indexes[q];
						break;
				}
			}
			a.nbr_type.push_back(type);
			a.orth_on_coarse.push_back(orth);
			a.nbr_idx.insert(a.nbr_idx.end(), idx, idx + Q);
		}
	}
	return a;
}
// the hierarchy for a list of the reference's domains (what GMG::CycleFactory walks, GMG/CycleFactory3d.cpp:86-121)
template <size_t D> std::shared_ptr<Handles> createHierarchy(const std::vector<std::shared_ptr<Domain<D>>> &domains, int n, int device = 0)
{
	auto hd = std::make_shared<Handles>();
	tg(tgpu_init(device, &hd->ctx));
	std::vector<LevelArrays<D>> arrays;
	std::vector<TgpuLevelDesc>  descs;
	for (size_t l = 0; l < domains.size(); l++) arrays.push_back(flatten<D>(*domains[l], l + 1 < domains.size() ? domains[l + 1].get() : nullptr));
	for (auto &a : arrays) descs.push_back(a.desc());
	tg(tgpu_hierarchy_create(hd->ctx, (int) D, n, (int) descs.size(), descs.data(), &hd->h));
	return hd;
}

// ---------------------------------------------------------------------------------------------------------------
// Vector<D> on the device.  Every op of Vector.h:190-321 is overridden with its tgpu_vec_* kernel.  getLocalData (the
// cold path: Init, Domain::integrate, writers, mixed operations with host vectors through the base-class loops) hands
// out views of a host mirror kept coherent with the device copy: the LocalDataManager hook (Vector.h:30-34) marks the
// mirror dirty when a writable view is released, device ops flush a dirty mirror first (one H2D), and a view is only
// re-downloaded when a device op has run since the last download.
// ---------------------------------------------------------------------------------------------------------------
template <size_t D> class TgpuVector : public Vector<D>
{
	std::shared_ptr<Handles>    hd;
	tgpu_vec *                  v = nullptr;
	std::array<int, D>          ns, strides;
	int                         patch_stride = 1;
	mutable std::vector<double> mirror;
	mutable bool                mirror_valid = false; // mirror == device copy (or newer, if dirty)
	mutable bool                mirror_dirty = false; // host wrote through a view; the device copy is stale

	struct Release : LocalDataManager {
		const TgpuVector *o;
		bool              writable;
		Release(const TgpuVector *o, bool w) : o(o), writable(w) {}
		~Release()
		{
			if (writable) o->mirror_dirty = true;
		}
	};
	void acquire() const
	{
		if (mirror_valid) return;
		mirror.resize((size_t) this->num_local_patches * patch_stride);
		tg(tgpu_vec_download(v, mirror.data()));
		mirror_valid = true;
	}

	public:
	TgpuVector(std::shared_ptr<Handles> hd, int level, int n) : hd(hd)
	{
		tg(tgpu_vec_create(hd->h, level, &v));
		int64_t np = 0, nc = 0;
		tg(tgpu_level_npatch(hd->h, level, &np, &nc));
		this->num_local_patches = (int) np;
		ns.fill(n);
		for (size_t i = 0; i < D; i++) {
			strides[i] = patch_stride;
			patch_stride *= n;
		}
	}
	~TgpuVector() { tgpu_vec_destroy(v); }
	TgpuVector(const TgpuVector &) = delete;
	TgpuVector &operator=(const TgpuVector &) = delete;

	// device handle for an ABI call that READS the vector: a dirty mirror goes to the device first
	const tgpu_vec *dev() const
	{
		if (mirror_dirty) {
			tg(tgpu_vec_upload(v, mirror.data()));
			mirror_dirty = false;
		}
		return v;
	}
	// device handle for an ABI call that WRITES the vector: the mirror is stale afterwards
	tgpu_vec *dev_mut()
	{
		dev();
		mirror_valid = false;
		return v;
	}
	static const TgpuVector *cast(const Vector<D> *b)
	{
		auto d = dynamic_cast<const TgpuVector *>(b);
		if (!d) throw 3;
		return d;
	}
	static TgpuVector *cast(Vector<D> *b)
	{
		auto d = dynamic_cast<TgpuVector *>(b);
		if (!d) throw 3;
		return d;
	}
	static const tgpu_vec *in(const std::shared_ptr<const Vector<D>> &b) { return cast(b.get())->dev(); }
	static tgpu_vec *      out(const std::shared_ptr<Vector<D>> &b) { return cast(b.get())->dev_mut(); }
	static bool            is(const std::shared_ptr<const Vector<D>> &b) { return dynamic_cast<const TgpuVector *>(b.get()) != nullptr; }

	LocalData<D> getLocalData(int i) override
	{
		acquire();
		return LocalData<D>(mirror.data() + (size_t) i * patch_stride, strides, ns, std::make_shared<Release>(this, true));
	}
	const LocalData<D> getLocalData(int i) const override
	{
		acquire();
		return LocalData<D>(mirror.data() + (size_t) i * patch_stride, strides, ns, std::make_shared<Release>(this, false));
	}

	void set(double a) override { tg(tgpu_vec_set(dev_mut(), a)); }
	void scale(double a) override { tg(tgpu_vec_scale(dev_mut(), a)); }
	void shift(double d) override { tg(tgpu_vec_shift(dev_mut(), d)); }
	// operands that are not device vectors (e.g. a PetscVector filled by Init) take the reference's own loops over
	// getLocalData (Vector.h:215-262), i.e. the host mirror
	void copy(std::shared_ptr<const Vector<D>> b) override
	{
		if (is(b)) tg(tgpu_vec_copy(dev_mut(), in(b)));
		else Vector<D>::copy(b);
	}
	void add(std::shared_ptr<const Vector<D>> b) override
	{
		if (is(b)) tg(tgpu_vec_add(dev_mut(), in(b)));
		else Vector<D>::add(b);
	}
	void addScaled(double a, std::shared_ptr<const Vector<D>> b) override
	{
		if (is(b)) tg(tgpu_vec_add_scaled(dev_mut(), a, in(b)));
		else Vector<D>::addScaled(a, b);
	}
	void addScaled(double a, std::shared_ptr<const Vector<D>> x, double b, std::shared_ptr<const Vector<D>> y) override
	{
		if (is(x) && is(y)) tg(tgpu_vec_add_scaled2(dev_mut(), a, in(x), b, in(y)));
		else Vector<D>::addScaled(a, x, b, y);
	}
	void scaleThenAdd(double a, std::shared_ptr<const Vector<D>> b) override
	{
		if (is(b)) tg(tgpu_vec_scale_then_add(dev_mut(), a, in(b)));
		else Vector<D>::scaleThenAdd(a, b);
	}
	void scaleThenAddScaled(double a, double b, std::shared_ptr<const Vector<D>> x) override
	{
		if (is(x)) tg(tgpu_vec_scale_then_add_scaled(dev_mut(), a, b, in(x)));
		else Vector<D>::scaleThenAddScaled(a, b, x);
	}
	void scaleThenAddScaled(double a, double b, std::shared_ptr<const Vector<D>> x, double g, std::shared_ptr<const Vector<D>> y) override
	{
		if (is(x) && is(y)) tg(tgpu_vec_scale_then_add_scaled2(dev_mut(), a, b, in(x), g, in(y)));
		else Vector<D>::scaleThenAddScaled(a, b, x, g, y);
	}
	double twoNorm() const override
	{
		double r;
		tg(tgpu_vec_two_norm(dev(), &r));
		return r;
	}
	double infNorm() const override
	{
		double r;
		tg(tgpu_vec_inf_norm(dev(), &r));
		return r;
	}
	double dot(std::shared_ptr<const Vector<D>> b) const override
	{
		if (!is(b)) return Vector<D>::dot(b);
		double r;
		tg(tgpu_vec_dot(dev(), in(b), &r));
		return r;
	}
};

template <size_t D> class TgpuVG : public VectorGenerator<D>
{
	std::shared_ptr<Handles> hd;
	int                      level, n;

	public:
	TgpuVG(std::shared_ptr<Handles> hd, int level, int n) : hd(hd), level(level), n(n) {}
	std::shared_ptr<Vector<D>> getNewVector() override { return std::make_shared<TgpuVector<D>>(hd, level, n); }
};

template <size_t D> class TgpuOp : public Operator<D>
{
	std::shared_ptr<Handles> hd;
	int                      level;

	public:
	TgpuOp(std::shared_ptr<Handles> hd, int level) : hd(hd), level(level) {}
	void apply(std::shared_ptr<const Vector<D>> x, std::shared_ptr<Vector<D>> b) const override
	{
		tg(tgpu_apply(hd->h, level, TgpuVector<D>::in(x), TgpuVector<D>::out(b)));
	}
};
template <size_t D> class TgpuSmoother : public GMG::Smoother<D>
{
	std::shared_ptr<Handles> hd;
	int                      level;

	public:
	TgpuSmoother(std::shared_ptr<Handles> hd, int level) : hd(hd), level(level) {}
	void smooth(std::shared_ptr<const Vector<D>> f, std::shared_ptr<Vector<D>> u) const override
	{
		tg(tgpu_smooth(hd->h, level, TgpuVector<D>::in(f), TgpuVector<D>::out(u)));
	}
};
// the north star's weighted-Jacobi option (no reference counterpart), usable wherever a GMG::Smoother<D> is
template <size_t D> class TgpuJacobiSmoother : public GMG::Smoother<D>
{
	std::shared_ptr<Handles> hd;
	int                      level;
	double                   omega;

	public:
	TgpuJacobiSmoother(std::shared_ptr<Handles> hd, int level, double omega) : hd(hd), level(level), omega(omega) {}
	void smooth(std::shared_ptr<const Vector<D>> f, std::shared_ptr<Vector<D>> u) const override
	{
		tg(tgpu_smooth_jacobi(hd->h, level, TgpuVector<D>::in(f), TgpuVector<D>::out(u), omega));
	}
};
template <size_t D> class TgpuRestrictor : public GMG::Restrictor<D>
{
	std::shared_ptr<Handles> hd;
	int                      fine_level;

	public:
	TgpuRestrictor(std::shared_ptr<Handles> hd, int fine_level) : hd(hd), fine_level(fine_level) {}
	void restrict(std::shared_ptr<Vector<D>> coarse, std::shared_ptr<const Vector<D>> fine) const override
	{
		tg(tgpu_restrict(hd->h, fine_level, TgpuVector<D>::in(fine), TgpuVector<D>::out(coarse)));
	}
};
template <size_t D> class TgpuInterpolator : public GMG::Interpolator<D>
{
	std::shared_ptr<Handles> hd;
	int                      fine_level;
	bool                     linear;

	public:
	TgpuInterpolator(std::shared_ptr<Handles> hd, int fine_level, bool linear = false) : hd(hd), fine_level(fine_level), linear(linear) {}
	void interpolate(std::shared_ptr<const Vector<D>> coarse, std::shared_ptr<Vector<D>> fine) const override
	{
		if (linear) tg(tgpu_prolong_add_linear(hd->h, fine_level, TgpuVector<D>::in(coarse), TgpuVector<D>::out(fine)));
		else tg(tgpu_prolong_add(hd->h, fine_level, TgpuVector<D>::in(coarse), TgpuVector<D>::out(fine)));
	}
};

inline TgpuCycleOpts toAbi(const GMG::CycleOpts &opts)
{
	TgpuCycleOpts o;
	tgpu_cycle_opts_default(&o);
	o.max_levels       = opts.max_levels;
	o.patches_per_proc = opts.patches_per_proc;
	o.pre_sweeps       = opts.pre_sweeps;
	o.post_sweeps      = opts.post_sweeps;
	o.mid_sweeps       = opts.mid_sweeps;
	o.coarse_sweeps    = opts.coarse_sweeps;
	if (opts.cycle_type == "V") o.cycle_type = 0;
	else if (opts.cycle_type == "W") o.cycle_type = 1;
	else throw 3; // GMG/CycleFactory3d.cpp:131
	return o;
}
// The whole cycle as ONE ABI call (fused kernel schedule, CUDA-graph replay).  Is-an Operator<D> exactly like
// GMG::Cycle<D> (GMG/Cycle.h:34), so BiCGStab (BiCGStab.h:57,73-83) and PetscShellCreator (PetscShellCreator.h:44-75)
// take it unchanged as the preconditioner.
template <size_t D> class TgpuCycle : public Operator<D>
{
	std::shared_ptr<Handles> hd;
	TgpuCycleOpts            o;

	public:
	TgpuCycle(std::shared_ptr<Handles> hd, const GMG::CycleOpts &opts) : hd(hd), o(toAbi(opts)) {}
	void apply(std::shared_ptr<const Vector<D>> f, std::shared_ptr<Vector<D>> u) const override
	{
		tg(tgpu_vcycle(hd->h, &o, TgpuVector<D>::in(f), TgpuVector<D>::out(u)));
	}
};

// GMG::CycleFactory3d::getCycle's level list (GMG/CycleFactory3d.cpp:86-121) over the adaptors: one reference
// GMG::Level<D> per hierarchy level with operator, smoother, restrictor and interpolator set and the levels linked.
// `global_patches[l]` = Domain::getNumGlobalPatches() of level l, for the patches_per_proc rule (:104).
template <size_t D>
std::shared_ptr<GMG::Level<D>> getLevels(std::shared_ptr<Handles> hd, int n, const GMG::CycleOpts &opts, const std::vector<int> &global_patches, int nranks = 1)
{
	std::shared_ptr<GMG::Level<D>> finest, finer;
	for (int l = 0; l < (int) global_patches.size(); l++) {
		if (l > 0 && opts.max_levels > 0 && l >= opts.max_levels) break;
		if (l > 0 && (global_patches[l] + 0.0) / nranks < opts.patches_per_proc) break;
		std::shared_ptr<VectorGenerator<D>> vg(new TgpuVG<D>(hd, l, n));
		std::shared_ptr<GMG::Level<D>>      level(new GMG::Level<D>(vg));
		level->setOperator(std::shared_ptr<Operator<D>>(new TgpuOp<D>(hd, l)));
		level->setSmoother(std::shared_ptr<GMG::Smoother<D>>(new TgpuSmoother<D>(hd, l)));
		if (finer) {
			level->setFiner(finer);
			finer->setCoarser(level);
			finer->setRestrictor(std::shared_ptr<GMG::Restrictor<D>>(new TgpuRestrictor<D>(hd, l - 1)));
			level->setInterpolator(std::shared_ptr<GMG::Interpolator<D>>(new TgpuInterpolator<D>(hd, l - 1)));
		} else {
			finest = level;
		}
		finer = level;
	}
	return finest;
}
// the reference's own GMG::VCycle / GMG::WCycle objects driving the adaptors, one ABI call per smoother / operator /
// transfer step (the plugin-granular path); TgpuCycle above is the one-call form of the same cycle
template <size_t D>
std::shared_ptr<GMG::Cycle<D>> getCycle(std::shared_ptr<Handles> hd, int n, const GMG::CycleOpts &opts, const std::vector<int> &global_patches, int nranks = 1)
{
	auto finest = getLevels<D>(hd, n, opts, global_patches, nranks);
	if (opts.cycle_type == "V") return std::shared_ptr<GMG::Cycle<D>>(new GMG::VCycle<D>(finest, opts));
	if (opts.cycle_type == "W") return std::shared_ptr<GMG::Cycle<D>>(new GMG::WCycle<D>(finest, opts));
	throw 3;
}
} // namespace tgpu_te
