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
	{
		// the record count must agree with the file size for this D (a 3D file opened with D = 2, or a truncated
		// file, is rejected before any record is interpreted)
		const std::streamoff rec = 12 + 16 * D + 8 * D + 4 * (1 << D);
		in.seekg(0, std::ios::end);
		const std::streamoff size = in.tellg();
		in.seekg(8, std::ios::beg);
		if (size != 8 + rec * (std::streamoff) hdr[0])
			throw std::runtime_error("mesh: " + path + " does not hold " + std::to_string(hdr[0]) + " records of a " + std::to_string(D)
			                         + "D tree (wrong D or truncated file)");
	}
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
		if (n.level < 1 || n.level > 30) throw std::runtime_error("mesh: node level out of range in " + path);
		if (n.id >= 0 && (size_t) n.id < m.nodes.size() && m.nodes[n.id].id >= 0) throw std::runtime_error("mesh: duplicate node id in " + path);
		m.put(n);
		m.num_levels = std::max(m.num_levels, n.level);
	}
	m.validate(path);
	return m;
}

// Every id stored in a record is used as a raw index later (refineNode, extractLevels): check them all once, and the
// parent/child/neighbour back-links with them, so that a corrupt file is an I/O error instead of an out-of-bounds access
// (the reference keeps nodes in a std::map and cannot index out of range).
void Mesh::validate(const std::string &path) const
{
	auto bad = [&](const char *what, int id) { throw std::runtime_error("mesh: " + std::string(what) + " (node " + std::to_string(id) + ") in " + path); };
	auto ok  = [&](int id) { return id == -1 || (id >= 0 && (size_t) id < nodes.size() && nodes[id].id == id); };
	const int no = 1 << D;
	for (const MeshNode &n : nodes) {
		if (n.id < 0) continue;
		if (!ok(n.parent)) bad("dangling parent id", n.id);
		for (int a = 0; a < D; a++)
			if (!(n.lengths[a] > 0.0)) bad("non-positive patch length", n.id);
		for (int s = 0; s < 2 * D; s++) {
			if (!ok(n.nbr[s])) bad("dangling neighbour id", n.id);
			if (n.nbr[s] != -1) {
				const MeshNode &b = nodes[n.nbr[s]];
				if (b.nbr[s ^ 1] != n.id) bad("neighbour link is not mutual", n.id);
				if (b.level != n.level) bad("neighbour on a different tree level", n.id);
			}
		}
		int nchild = 0;
		for (int o = 0; o < no; o++) {
			if (!ok(n.child[o])) bad("dangling child id", n.id);
			if (n.child[o] != -1) {
				nchild++;
				const MeshNode &c = nodes[n.child[o]];
				if (c.parent != n.id || c.level != n.level + 1) bad("child does not point back to its parent", n.id);
			}
		}
		if (nchild != 0 && nchild != no) bad("partially refined node", n.id);
		if (n.parent != -1) {
			const MeshNode &p     = nodes[n.parent];
			bool            found = false;
			for (int o = 0; o < no; o++) found = found || p.child[o] == n.id;
			if (!found) bad("parent does not list the node as a child", n.id);
		} else if (n.id != root) {
			bad("second root node", n.id);
		}
	}
	if (root < 0 || nodes[root].parent != -1) bad("first record is not the root", root);
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

void Mesh::refineBox(const double *lo, const double *hi)
{
	std::vector<std::pair<int, int>> leaves;
	int                              deepest = 0;
	for (const MeshNode &n : nodes) {
		if (n.id < 0 || n.hasChildren()) continue;
		bool in = true;
		for (int a = 0; a < D; a++) {
			const double c = n.starts[a] + 0.5 * n.lengths[a];
			in             = in && c >= lo[a] && c < hi[a];
		}
		if (in) leaves.push_back({n.level, n.id}), deepest = std::max(deepest, n.level);
	}
	std::sort(leaves.begin(), leaves.end());
	nodes.reserve(nodes.size() + leaves.size() * (size_t) (1 << D));
	for (auto &p : leaves) refineNode(p.second);
	if (!leaves.empty()) num_levels = std::max(num_levels, deepest + 1);
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
					while (quad < Q && parent.child[octs[quad]] != nd.id) quad++;
					if (quad == Q) throw std::runtime_error("mesh: node without a sibling link on an interior side (corrupt tree)");
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
					while (o < (1 << D) && nodes[nd.parent].child[o] != nd.id) o++;
					if (o == (1 << D)) throw std::runtime_error("mesh: parent does not list the node as a child");
					L.orth_on_parent[k] = (int8_t) o;
				}
			}
			for (int s = 0; s < S; s++) {
				L.nbr_type[(size_t) k * S + s]       = r.type[s];
				// ThundereggDomGen(..., neumann = true): every side on the domain boundary (ThundereggDomGen.h:216-220,
				// PatchInfo::setNeumann PatchInfo.h:684-697 with an always-true predicate)
				if (neumann && r.type[s] == TGPU_NBR_NONE) L.neumann[k] |= (uint8_t) (1u << s);
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

// ---------------------------------------------------------------------------------------------
// partition + halo plan
// ---------------------------------------------------------------------------------------------
namespace tgpu
{
static uint64_t mortonKey(const double *starts, int D)
{
	uint64_t key = 0;
	uint32_t c[3] = {0, 0, 0};
	for (int a = 0; a < D; a++) {
		double s = starts[a];
		if (s < 0) s = 0;
		if (s >= 1) s = 0.999999999;
		c[a] = (uint32_t) (s * (double) (1u << 20));
	}
	for (int b = 19; b >= 0; b--)
		for (int a = D - 1; a >= 0; a--) key = (key << 1) | ((c[a] >> b) & 1u);
	return key;
}

Partition partitionLevels(const std::vector<HostLevel> &global, int D, int n, int rank, int nranks, int min_patches_per_rank)
{
	if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("partition: bad rank / nranks");
	const int L = (int) global.size(), S = 2 * D, Q = 1 << (D - 1);
	Partition part;
	part.D = D, part.n = n, part.rank = rank, part.nranks = nranks;
	// distributed levels = the finest levels that have enough patches for every rank
	int ndist = 0;
	if (nranks > 1)
		while (ndist < L && global[ndist].npatch >= (int64_t) nranks * std::max(1, min_patches_per_rank)) ndist++;
	part.ndist = ndist;
	part.owner.resize(ndist);
	if (ndist > 0) {
		// weights: patches visited per cycle below (and including) each patch of the coarsest distributed
		// level, summed over ALL distributed levels - a coarse leaf that persists through k levels is
		// smoothed k times, so it weighs k
		std::vector<std::vector<int64_t>> w(ndist);
		w[0].assign(global[0].npatch, 1);
		for (int l = 0; l + 1 < ndist; l++) {
			w[l + 1].assign(global[l + 1].npatch, 1);
			for (int p = 0; p < global[l].npatch; p++) w[l + 1][global[l].parent_idx[p]] += w[l][p];
		}
		const int            top = ndist - 1;
		const HostLevel &    T   = global[top];
		std::vector<int32_t> order(T.npatch);
		for (int p = 0; p < T.npatch; p++) order[p] = p;
		std::vector<uint64_t> key(T.npatch);
		for (int p = 0; p < T.npatch; p++) key[p] = mortonKey(&T.starts[(size_t) p * D], D);
		std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
		int64_t total = 0;
		for (int64_t x : w[top]) total += x;
		part.owner[top].assign(T.npatch, 0);
		// contiguous Morton ranges of ~equal weight; every rank gets at least one patch
		int64_t acc = 0;
		int     r = 0, have = 0;
		for (int i = 0; i < T.npatch; i++) {
			const int32_t p    = order[i];
			part.owner[top][p] = r;
			acc += w[top][p];
			have++;
			const bool quota_met   = acc * nranks >= (int64_t) (r + 1) * total;
			const bool must_advance = (T.npatch - i - 1) <= (nranks - r - 1);
			if (r < nranks - 1 && have >= 1 && (quota_met || must_advance)) r++, have = 0;
		}
		for (int l = top - 1; l >= 0; l--) {
			part.owner[l].resize(global[l].npatch);
			for (int p = 0; p < global[l].npatch; p++) part.owner[l][p] = part.owner[l + 1][global[l].parent_idx[p]];
		}
	}
	part.levels.resize(L);
	std::vector<int32_t> local_of_global_next; // local index of level l+1 patches on this rank (distributed) or identity
	for (int l = L - 1; l >= 0; l--) {
		PartLevel &      PL = part.levels[l];
		const HostLevel &GL = global[l];
		std::vector<int32_t> local_of_global(GL.npatch, -1);
		if (l >= ndist) {
			PL.distributed = false;
			PL.local       = GL;
			PL.n_owned     = GL.npatch;
			PL.n_interior  = GL.npatch;
			PL.owned_global.resize(GL.npatch);
			for (int p = 0; p < GL.npatch; p++) PL.owned_global[p] = p, local_of_global[p] = p;
			local_of_global_next = local_of_global;
			continue;
		}
		PL.distributed                 = true;
		const std::vector<int32_t> &ow = part.owner[l];
		// owned patches: interior ones (no off-rank neighbour) first, then the boundary ones, so that the
		// interior sweep can run while the halo exchange is in flight
		std::vector<int32_t> interior, boundary;
		for (int p = 0; p < GL.npatch; p++) {
			if (ow[p] != rank) continue;
			bool bnd = false;
			for (int s = 0; s < S && !bnd; s++)
				for (int q = 0; q < Q; q++) {
					const int32_t j = GL.nbr_idx[((size_t) p * S + s) * Q + q];
					if (j >= 0 && ow[j] != rank) bnd = true;
				}
			(bnd ? boundary : interior).push_back(p);
		}
		PL.n_interior = (int32_t) interior.size();
		PL.owned_global = interior;
		PL.owned_global.insert(PL.owned_global.end(), boundary.begin(), boundary.end());
		for (size_t k = 0; k < PL.owned_global.size(); k++) local_of_global[PL.owned_global[k]] = (int32_t) k;
		PL.n_owned = (int32_t) PL.owned_global.size();
		// halo = off-rank neighbours of owned patches, in global order
		std::vector<char> is_halo(GL.npatch, 0);
		for (int32_t p : PL.owned_global)
			for (int s = 0; s < S; s++)
				for (int q = 0; q < Q; q++) {
					const int32_t j = GL.nbr_idx[((size_t) p * S + s) * Q + q];
					if (j >= 0 && ow[j] != rank) is_halo[j] = 1;
				}
		for (int p = 0; p < GL.npatch; p++)
			if (is_halo[p]) {
				local_of_global[p] = PL.n_owned + (int32_t) PL.halo_global.size();
				PL.halo_global.push_back(p);
				PL.halo_owner.push_back(ow[p]);
			}
		PL.n_halo = (int32_t) PL.halo_global.size();
		// remapped local tables
		HostLevel &LL = PL.local;
		const int  P  = PL.n_owned + PL.n_halo;
		LL.npatch     = P;
		LL.spacing.resize((size_t) P * D);
		LL.starts.resize((size_t) P * D);
		LL.neumann.assign(P, 0);
		LL.nbr_type.assign((size_t) P * S, TGPU_NBR_NONE);
		LL.orth_on_coarse.assign((size_t) P * S, -1);
		LL.orth_on_parent.assign(P, -1);
		LL.nbr_idx.assign((size_t) P * S * Q, -1);
		LL.parent_idx.assign(P, -1);
		LL.ids.resize(P);
		LL.parent_ids.resize(P);
		LL.refine_levels.resize(P);
		for (int k = 0; k < P; k++) {
			const int32_t gp = k < PL.n_owned ? PL.owned_global[k] : PL.halo_global[k - PL.n_owned];
			for (int a = 0; a < D; a++) {
				LL.spacing[(size_t) k * D + a] = GL.spacing[(size_t) gp * D + a];
				LL.starts[(size_t) k * D + a]  = GL.starts[(size_t) gp * D + a];
			}
			LL.ids[k]           = GL.ids[gp];
			LL.parent_ids[k]    = GL.parent_ids[gp];
			LL.refine_levels[k] = GL.refine_levels[gp];
			if (k >= PL.n_owned) continue; // halo slots carry faces only: no neighbours, no parent
			LL.neumann[k]        = GL.neumann[gp];
			LL.orth_on_parent[k] = GL.orth_on_parent[gp];
			if (l + 1 < L) {
				const int32_t pl = local_of_global_next[GL.parent_idx[gp]];
				if (pl < 0) throw std::runtime_error("partition: parent of an owned patch is not local");
				LL.parent_idx[k] = pl;
			}
			for (int s = 0; s < S; s++) {
				LL.nbr_type[(size_t) k * S + s]       = GL.nbr_type[(size_t) gp * S + s];
				LL.orth_on_coarse[(size_t) k * S + s] = GL.orth_on_coarse[(size_t) gp * S + s];
				for (int q = 0; q < Q; q++) {
					const int32_t j = GL.nbr_idx[((size_t) gp * S + s) * Q + q];
					if (j >= 0) LL.nbr_idx[((size_t) k * S + s) * Q + q] = local_of_global[j];
				}
			}
		}
		// exchange plan: peer q needs face (N, s^1) of my patch N whenever one of q's patches has N on side s
		struct Need {
			int32_t patch, side;
			bool    operator<(const Need &o) const { return patch != o.patch ? patch < o.patch : side < o.side; }
			bool    operator==(const Need &o) const { return patch == o.patch && side == o.side; }
		};
		std::vector<std::vector<Need>> send(nranks), recv(nranks);
		for (int p = 0; p < GL.npatch; p++)
			for (int s = 0; s < S; s++)
				for (int q = 0; q < Q; q++) {
					const int32_t j = GL.nbr_idx[((size_t) p * S + s) * Q + q];
					if (j < 0 || ow[j] == ow[p]) continue;
					if (ow[p] == rank) recv[ow[j]].push_back({j, s ^ 1}); // I read face s^1 of j (owned by ow[j])
					if (ow[j] == rank) send[ow[p]].push_back({j, s ^ 1}); // ow[p] reads face s^1 of my patch j
				}
		for (int r = 0; r < nranks; r++) {
			if (r == rank) continue;
			auto uniq = [](std::vector<Need> &v) {
				std::sort(v.begin(), v.end());
				v.erase(std::unique(v.begin(), v.end()), v.end());
			};
			uniq(send[r]);
			uniq(recv[r]);
			if (send[r].empty() && recv[r].empty()) continue;
			PeerExchange px;
			px.peer = r;
			for (const Need &x : send[r]) {
				px.send_patch.push_back(local_of_global[x.patch]);
				px.send_side.push_back(x.side);
			}
			for (const Need &x : recv[r]) {
				px.recv_slot.push_back(local_of_global[x.patch]);
				px.recv_side.push_back(x.side);
			}
			PL.peers.push_back(std::move(px));
		}
		local_of_global_next = local_of_global;
	}
	return part;
}
} // namespace tgpu
