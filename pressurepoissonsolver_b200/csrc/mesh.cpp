// mesh.cpp - see mesh.h.  Host only; no CUDA.
#include "mesh.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <stdexcept>

namespace tgpu
{
void orthantsOnSide(int D, int side, int out[4])
{
	const int bit = side / 2, set = side & 1;
	for (int i = 0; i < (1 << (D - 1)); i++) {
		const int lower = i & ((1 << bit) - 1);
		const int upper = (i >> bit) << (bit + 1);
		out[i]          = upper | lower | (set << bit);
	}
}

TgpuLevelDesc HostLevel::desc() const
{
	TgpuLevelDesc d;
	d.npatch         = npatch;
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

void Mesh::put(const MeshNode &n)
{
	if (n.id < 0 || n.id > (1 << 28)) throw std::runtime_error("mesh: node id out of range");
	if ((size_t) n.id >= nodes.size()) nodes.resize((size_t) n.id + 1);
	nodes[n.id] = n;
	max_id      = std::max(max_id, n.id);
}

int Mesh::numNodes() const
{
	int c = 0;
	for (const MeshNode &n : nodes) c += n.id >= 0;
	return c;
}

// file format: SURVEY App. B / OctTree.h:90-118 - little-endian packed records
Mesh Mesh::load(const std::string &path, int D)
{
	if (D != 2 && D != 3) throw std::runtime_error("mesh: D must be 2 or 3");
	std::ifstream in(path, std::ios::binary);
	if (!in) throw std::runtime_error("mesh: cannot open " + path);
	Mesh m;
	m.D = D;
	int32_t hdr[2];
	in.read((char *) hdr, 8);
	if (!in || hdr[0] <= 0) throw std::runtime_error("mesh: bad header in " + path);
	for (int i = 0; i < hdr[0]; i++) {
		MeshNode n;
		int32_t  ilp[3];
		in.read((char *) ilp, 12);
		n.id = ilp[0], n.level = ilp[1], n.parent = ilp[2];
		in.read((char *) n.lengths, 8 * D);
		in.read((char *) n.starts, 8 * D);
		in.read((char *) n.nbr, 4 * 2 * D);
		in.read((char *) n.child, 4 * (1 << D));
		if (!in) throw std::runtime_error("mesh: truncated file " + path);
		if (i == 0) m.root = n.id;
		m.put(n);
		m.num_levels = std::max(m.num_levels, n.level);
	}
	return m;
}

Mesh Mesh::uniform(int D, int num_levels)
{
	if (D != 2 && D != 3) throw std::runtime_error("mesh: D must be 2 or 3");
	Mesh     m;
	MeshNode r;
	m.D     = D;
	r.id    = 0;
	r.level = 1; // shipped mesh files number the root level 1
	for (int i = 0; i < D; i++) r.lengths[i] = 1.0, r.starts[i] = 0.0;
	m.root       = 0;
	m.num_levels = 1;
	m.put(r);
	for (int l = 1; l < num_levels; l++) m.refineLeaves();
	return m;
}

// OctTree.h:119-179: every leaf is refined once, in (level, id) order (std::set<pair> iteration
// order of the BFS over face neighbours; the BFS reaches every leaf of a connected domain).
void Mesh::refineLeaves()
{
	std::vector<std::pair<int, int>> leaves;
	for (const MeshNode &n : nodes)
		if (n.id >= 0 && !n.hasChildren()) leaves.push_back({n.level, n.id});
	std::sort(leaves.begin(), leaves.end());
	nodes.reserve(nodes.size() + leaves.size() * (size_t) (1 << D));
	for (auto &p : leaves) refineNode(p.second);
	num_levels++;
}

// OctTree.h:180-213 + Node(parent, orthant) OctNode.h:78-90
void Mesh::refineNode(int id)
{
	const int no = 1 << D;
	MeshNode  kids[8];
	{
		MeshNode &n = nodes[id];
		for (int o = 0; o < no; o++) {
			MeshNode &c = kids[o];
			c.parent    = n.id;
			c.level     = n.level + 1;
			for (int i = 0; i < D; i++) {
				c.lengths[i] = n.lengths[i] / 2;
				c.starts[i]  = ((o >> i) & 1) ? n.starts[i] + c.lengths[i] : n.starts[i];
			}
			c.id       = ++max_id;
			n.child[o] = c.id;
		}
	}
	for (int o = 0; o < no; o++)
		for (int i = 0; i < D; i++) {
			const int s    = 2 * i + (((o >> i) & 1) ? 0 : 1); // interior side on axis i
			kids[o].nbr[s] = kids[o ^ (1 << i)].id;
		}
	nodes.resize((size_t) max_id + 1);
	for (int s = 0; s < 2 * D; s++) {
		const int nb = nodes[id].nbr[s];
		if (nb != -1 && nodes[nb].hasChildren()) {
			int octs[4];
			orthantsOnSide(D, s, octs);
			for (int q = 0; q < no / 2; q++) {
				MeshNode &child     = kids[octs[q]];
				MeshNode &nbr_child = nodes[nodes[nb].child[octs[q] ^ (1 << (s / 2))]];
				child.nbr[s]        = nbr_child.id;
				nbr_child.nbr[s ^ 1] = child.id;
			}
		}
	}
	for (int o = 0; o < no; o++) nodes[kids[o].id] = kids[o];
}

std::vector<HostLevel> Mesh::extractLevels(int n) const
{
	const int              Q = 1 << (D - 1), S = 2 * D;
	std::vector<HostLevel> out;
	std::vector<int32_t>   local_of_id(nodes.size(), -1), prev_local;
	for (int curr = num_levels; curr >= 1; curr--) {
		// members of level `curr`: nodes at tree level curr plus leaves above it
		// (ThundereggDomGen.h:127-222; the BFS there only establishes membership)
		std::vector<int32_t> members;
		for (const MeshNode &nd : nodes)
			if (nd.id >= 0 && (nd.level == curr || (nd.level < curr && !nd.hasChildren()))) members.push_back(nd.id);
		struct Rec {
			int8_t  type[6];
			int8_t  orth[6];
			int32_t ids[6][4];
		};
		std::vector<int32_t> rec_of_id(nodes.size(), -1);
		std::vector<Rec>     recs(members.size());
		for (size_t k = 0; k < members.size(); k++) {
			const MeshNode &nd = nodes[members[k]];
			Rec &           r  = recs[k];
			rec_of_id[nd.id]   = (int32_t) k;
			for (int s = 0; s < S; s++) {
				r.type[s] = TGPU_NBR_NONE;
				r.orth[s] = -1;
				for (int q = 0; q < 4; q++) r.ids[s][q] = -1;
				if (nd.nbr[s] == -1 && nd.parent != -1 && nodes[nd.parent].nbr[s] != -1) {
					const MeshNode &parent = nodes[nd.parent];
					int             octs[4], quad = 0;
					orthantsOnSide(D, s, octs);
					while (parent.child[octs[quad]] != nd.id) quad++;
					r.type[s]   = TGPU_NBR_COARSE;
					r.orth[s]   = (int8_t) quad;
					r.ids[s][0] = parent.nbr[s];
				} else if (nd.level < curr && nd.nbr[s] != -1 && nodes[nd.nbr[s]].hasChildren()) {
					int octs[4];
					orthantsOnSide(D, s ^ 1, octs);
					r.type[s] = TGPU_NBR_FINE;
					for (int q = 0; q < Q; q++) r.ids[s][q] = nodes[nd.nbr[s]].child[octs[q]];
				} else if (nd.nbr[s] != -1) {
					r.type[s]   = TGPU_NBR_NORMAL;
					r.ids[s][0] = nd.nbr[s];
				}
			}
		}
		// local order: BFS from the lowest id over getNbrIds() order (Domain.h:325-360)
		std::vector<int32_t> order;
		order.reserve(members.size());
		std::vector<char> enq(nodes.size(), 0);
		size_t            scan = 0; // members is ascending in id
		while (order.size() < members.size()) {
			while (enq[members[scan]]) scan++;
			std::deque<int32_t> bq;
			bq.push_back(members[scan]);
			enq[members[scan]] = 1;
			while (!bq.empty()) {
				const int32_t i = bq.front();
				bq.pop_front();
				order.push_back(i);
				const Rec &r = recs[rec_of_id[i]];
				for (int s = 0; s < S; s++)
					for (int q = 0; q < Q; q++) {
						const int32_t j = r.ids[s][q];
						if (j >= 0 && !enq[j]) {
							if (rec_of_id[j] < 0) throw std::runtime_error("mesh: neighbour outside level (unbalanced tree?)");
							enq[j] = 1;
							bq.push_back(j);
						}
					}
			}
		}
		std::fill(local_of_id.begin(), local_of_id.end(), -1);
		for (size_t k = 0; k < order.size(); k++) local_of_id[order[k]] = (int32_t) k;

		HostLevel L;
		const int P = (int) order.size();
		L.npatch    = P;
		L.spacing.resize((size_t) P * D);
		L.starts.resize((size_t) P * D);
		L.neumann.assign(P, 0);
		L.nbr_type.assign((size_t) P * S, TGPU_NBR_NONE);
		L.orth_on_coarse.assign((size_t) P * S, -1);
		L.orth_on_parent.assign(P, -1);
		L.nbr_idx.assign((size_t) P * S * Q, -1);
		L.parent_idx.assign(P, -1);
		L.ids.resize(P);
		L.parent_ids.resize(P);
		L.refine_levels.resize(P);
		for (int k = 0; k < P; k++) {
			const MeshNode &nd = nodes[order[k]];
			const Rec &     r  = recs[rec_of_id[nd.id]];
			L.ids[k]           = nd.id;
			L.refine_levels[k] = nd.level;
			for (int i = 0; i < D; i++) {
				L.spacing[(size_t) k * D + i] = nd.lengths[i] / n;
				L.starts[(size_t) k * D + i]  = nd.starts[i];
			}
			if (nd.level < curr) {
				L.parent_ids[k] = nd.id;
			} else {
				L.parent_ids[k] = nd.parent;
				if (nd.parent != -1) {
					int o = 0;
					while (nodes[nd.parent].child[o] != nd.id) o++;
					L.orth_on_parent[k] = (int8_t) o;
				}
			}
			for (int s = 0; s < S; s++) {
				L.nbr_type[(size_t) k * S + s]       = r.type[s];
				L.orth_on_coarse[(size_t) k * S + s] = r.orth[s];
				for (int q = 0; q < Q; q++)
					if (r.ids[s][q] >= 0) L.nbr_idx[((size_t) k * S + s) * Q + q] = local_of_id[r.ids[s][q]];
			}
		}
		out.push_back(std::move(L));
	}
	// parent local indices (what InterLevelComm resolves through AO maps, GMG/InterLevelComm.h:115-148)
	for (size_t l = 0; l + 1 < out.size(); l++) {
		std::vector<int32_t> loc(nodes.size(), -1);
		for (int k = 0; k < out[l + 1].npatch; k++) loc[out[l + 1].ids[k]] = k;
		for (int k = 0; k < out[l].npatch; k++) {
			const int32_t pid = out[l].parent_ids[k];
			if (pid < 0 || loc[pid] < 0) throw std::runtime_error("mesh: parent patch missing on the coarser level");
			out[l].parent_idx[k] = loc[pid];
		}
	}
	return out;
}
} // namespace tgpu
