// mesh.h - host-side octree/quadtree ingest and per-level patch metadata.
//
// Replaces, with flat arrays instead of std::map<int, Node> by-value copies, the reference's
//   Tree<D>(file) / refineLeaves / refineNode      src/Thunderegg/OctTree.h:90-213
//   ThundereggDomGen<D>::extractLevel               src/Thunderegg/ThundereggDomGen.h:127-222
//   Domain<D>::indexDomainsLocal (BFS local order)  src/Thunderegg/Domain.h:281-376
// The output is the TgpuLevelDesc list that tgpu_hierarchy_create turns into device tables.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tgpu.h"

namespace tgpu
{
struct MeshNode {
	int32_t id = -1, level = -1, parent = -1;
	double  lengths[3] = {0, 0, 0};
	double  starts[3]  = {0, 0, 0};
	int32_t nbr[6]     = {-1, -1, -1, -1, -1, -1};
	int32_t child[8]   = {-1, -1, -1, -1, -1, -1, -1, -1};
	bool hasChildren() const { return child[0] != -1; }
};

struct HostLevel {
	int32_t              npatch = 0;
	std::vector<double>  spacing, starts;
	std::vector<uint8_t> neumann;
	std::vector<int8_t>  nbr_type, orth_on_coarse, orth_on_parent;
	std::vector<int32_t> nbr_idx, parent_idx, ids, parent_ids, refine_levels;
	TgpuLevelDesc        desc() const;
};

class Mesh
{
	public:
	int D          = 3;
	int num_levels = 0;
	int root       = -1;
	int max_id     = -1;
	bool neumann   = false; // Neumann instead of Dirichlet conditions on the whole domain boundary

	std::vector<MeshNode> nodes; // indexed by id (ids are dense non-negative ints in every fixture)

	static Mesh load(const std::string &path, int D);
	static Mesh uniform(int D, int num_levels);
	void        refineLeaves();
	// refine every leaf whose centre lies in [lo, hi) (used to build weak-scaling meshes; the caller keeps 2:1 balance)
	void        refineBox(const double *lo, const double *hi);
	int         numNodes() const;

	// finest first; same patch order as the reference's local_index
	std::vector<HostLevel> extractLevels(int n) const;

	private:
	void      validate(const std::string &path) const;
	void      refineNode(int id);
	MeshNode &at(int id) { return nodes[id]; }
	void      put(const MeshNode &n);
};

// Orthant<D>::getValuesOnSide (src/Thunderegg/Side.h:346-362)
void orthantsOnSide(int D, int side, int out[4]);
} // namespace tgpu

// ---------------------------------------------------------------------------------------------
// Patch partition across the GPUs of one node + halo-exchange plan (replaces the reference's
// Zoltan PHG partition + PatchInfo migration, ThundereggDomGen.h:223-648, and the per-level
// VecScatter set-up of SchurHelper.h:195-280).
//
// Levels 0 .. ndist-1 are distributed: the coarsest distributed level is cut into `nranks`
// contiguous ranges along a Morton (Z-order) curve, weighted by the number of finest-level
// descendants; finer patches live where their parent lives, so restriction and prolongation
// between distributed levels are local.  Levels ndist .. are replicated on every rank (their
// right-hand side is assembled by an all-reduce).  A rank's local level = owned patches followed
// by halo slots for every off-rank neighbour; only face slices are exchanged.
namespace tgpu
{
struct PeerExchange {
	int                  peer = -1;
	std::vector<int32_t> send_patch, send_side; // local owned patch index, side whose face the peer needs
	std::vector<int32_t> recv_slot, recv_side;  // local halo slot index (>= n_owned), side
};
struct PartLevel {
	bool                      distributed = false;
	HostLevel                 local;        // remapped tables: npatch = n_owned + n_halo (replicated: the global level)
	int32_t                   n_owned = 0, n_halo = 0;
	int32_t                   n_interior = 0; // owned patches [0, n_interior) have no off-rank neighbour
	std::vector<int32_t>      owned_global; // global patch index of each owned patch
	std::vector<int32_t>      halo_global, halo_owner;
	std::vector<PeerExchange> peers;
};
struct Partition {
	int                    D = 3, n = 0, rank = 0, nranks = 1, ndist = 0;
	std::vector<PartLevel> levels;
	std::vector<std::vector<int32_t>> owner; // [level][global patch] for distributed levels
};
Partition partitionLevels(const std::vector<HostLevel> &global, int D, int n, int rank, int nranks, int min_patches_per_rank);
} // namespace tgpu
