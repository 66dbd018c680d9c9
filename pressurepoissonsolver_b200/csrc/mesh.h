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

	std::vector<MeshNode> nodes; // indexed by id (ids are dense non-negative ints in every fixture)

	static Mesh load(const std::string &path, int D);
	static Mesh uniform(int D, int num_levels);
	void        refineLeaves();
	int         numNodes() const;

	// finest first; same patch order as the reference's local_index
	std::vector<HostLevel> extractLevels(int n) const;

	private:
	void      refineNode(int id);
	MeshNode &at(int id) { return nodes[id]; }
	void      put(const MeshNode &n);
};

// Orthant<D>::getValuesOnSide (src/Thunderegg/Side.h:346-362)
void orthantsOnSide(int D, int side, int out[4]);
} // namespace tgpu
