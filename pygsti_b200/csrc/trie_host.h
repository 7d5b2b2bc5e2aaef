// trie_host.h -- host-side trie construction of the d = 16 path (plain C++, no CUDA: compiled into engine.cu and,
// stand-alone, into tests/trie_host_check.cpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <vector>

// ------------------------------------------------------------------------------------------------
// trie of circuits (prefix sharing) -- host side, once per atom.  Sequences are (root id, symbols...).
// The trie is built by scanning the circuits in lexicographic order, then cut into chains by a HEAVY-PATH
// decomposition: a chain runs from its head down through, at every node, the child with the deepest subtree; every
// other child starts a new chain.  A root-to-leaf path then changes chain O(log) times -- 8 hand-offs at most on the
// BASELINE layout instead of 25 (prefix trie) / 33 (suffix trie) when chains were "the nodes a circuit adds in scan
// order", and 40.8 k / 32.7 k chains instead of 68.3 k each -- which is what bounds k_trie_chains (dependency waits,
// not load latency: profiles/README.md).  Node ids are consecutive along a chain; a chain's parent node belongs to a
// chain with a smaller start depth (kernels_d16_trie.cuh relies on both).
// ------------------------------------------------------------------------------------------------
struct TrieHost {
    std::vector<int32_t> chain_parent; std::vector<uint32_t> chain_first, chain_len, chain_depth;
    std::vector<uint8_t> node_op;
    std::vector<uint32_t> depth_node;   // per circuit c, depth d in [0, L_c]: node id, at offset dptr[c] + d
    std::vector<uint64_t> dptr;
    std::vector<int64_t> sorted;        // circuits in lexicographic key order
};
static void build_trie(int64_t n, const std::vector<int32_t>& root, const std::vector<uint32_t>& ptr,
                       const std::vector<int32_t>& sym, bool reversed, TrieHost& T) {
    auto at = [&](int64_t c, uint32_t d) -> int32_t {      // d-th symbol of circuit c's key
        const uint32_t L = ptr[c + 1] - ptr[c];
        return reversed ? sym[ptr[c] + (L - 1 - d)] : sym[ptr[c] + d];
    };
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
        if (root[x] != root[y]) return root[x] < root[y];
        const uint32_t lx = ptr[x + 1] - ptr[x], ly = ptr[y + 1] - ptr[y];
        const uint32_t l = std::min(lx, ly);
        for (uint32_t d = 0; d < l; ++d) { const int32_t a = at(x, d), b = at(y, d); if (a != b) return a < b; }
        return lx < ly; });
    T.sorted = order;
    T.dptr.assign((size_t)n + 1, 0);
    for (int64_t c = 0; c < n; ++c) T.dptr[c + 1] = T.dptr[c] + (ptr[c + 1] - ptr[c]) + 1;
    T.depth_node.assign((size_t)T.dptr[n], 0);
    // ---- pass 1: the trie itself, temporary ids in creation order (a parent's id is smaller than its children's) ----
    const uint32_t NONE = 0xffffffffu;
    std::vector<uint32_t> tparent, tdepth; std::vector<uint8_t> top; std::vector<int32_t> troot;
    std::vector<uint32_t> path;
    int64_t prev = -1;
    for (int64_t oi = 0; oi < n; ++oi) {
        const int64_t c = order[oi];
        const uint32_t L = ptr[c + 1] - ptr[c];
        uint32_t lcp = 0;
        if (prev < 0 || root[prev] != root[c]) {
            const uint32_t id = (uint32_t)top.size();
            tparent.push_back(NONE); tdepth.push_back(0); top.push_back(255); troot.push_back(root[c]);
            path.assign(1, id);
        } else {
            const uint32_t Lp = ptr[prev + 1] - ptr[prev];
            const uint32_t l = std::min(L, Lp);
            while (lcp < l && at(prev, lcp) == at(c, lcp)) ++lcp;
            path.resize((size_t)lcp + 1);
        }
        for (uint32_t d = lcp; d < L; ++d) {
            const uint32_t id = (uint32_t)top.size();
            tparent.push_back(path[d]); tdepth.push_back(d + 1); top.push_back((uint8_t)at(c, d)); troot.push_back(0);
            path.push_back(id);
        }
        for (uint32_t d = 0; d <= L; ++d) T.depth_node[T.dptr[c] + d] = path[d];
        prev = c;
    }
    const size_t N = top.size();
    // ---- pass 2: subtree heights, heavy child (deepest subtree; the first created on ties) ----
    std::vector<uint32_t> height(N, 0), heavy(N, NONE);
    const bool lex_only = getenv("B200_TRIE_LEX") != nullptr;      // dev knob: continue into the first child in scan order instead
    for (size_t i = N; i-- > 0;) if (tparent[i] != NONE) height[tparent[i]] = std::max(height[tparent[i]], height[i] + 1);
    for (size_t i = 0; i < N; ++i) {
        const uint32_t p = tparent[i];
        if (p != NONE && (heavy[p] == NONE || (!lex_only && height[i] > height[heavy[p]]))) heavy[p] = (uint32_t)i;
    }
    // ---- pass 3: lay the chains out (heads in creation order: a head's parent lies on a chain laid out earlier) ----
    std::vector<uint32_t> newid(N, NONE);
    T.node_op.assign(N, 0);
    uint32_t next = 0;
    for (size_t h = 0; h < N; ++h) {
        const uint32_t p = tparent[h];
        if (p != NONE && heavy[p] == (uint32_t)h) continue;           // continues its parent's chain
        const uint32_t first = next;
        for (uint32_t v = (uint32_t)h; v != NONE; v = heavy[v]) { newid[v] = next; T.node_op[next] = top[v]; ++next; }
        T.chain_parent.push_back(p == NONE ? -(1 + troot[h]) : (int32_t)newid[p]);
        T.chain_first.push_back(first); T.chain_len.push_back(next - first); T.chain_depth.push_back(tdepth[h]);
    }
    for (auto& v : T.depth_node) v = newid[v];
    // work order of the chains: by start depth (a chain's parent chain starts at a smaller depth => is handed out
    // earlier: the spin-waits in k_trie_chains cannot deadlock), long chains first within a depth
    std::vector<uint32_t> ord(T.chain_first.size());
    std::iota(ord.begin(), ord.end(), 0u);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {
        if (T.chain_depth[x] != T.chain_depth[y]) return T.chain_depth[x] < T.chain_depth[y];
        return T.chain_len[x] > T.chain_len[y]; });
    std::vector<int32_t> cp(ord.size()); std::vector<uint32_t> cf(ord.size()), cl(ord.size());
    for (size_t i = 0; i < ord.size(); ++i) { cp[i] = T.chain_parent[ord[i]]; cf[i] = T.chain_first[ord[i]]; cl[i] = T.chain_len[ord[i]]; }
    T.chain_parent.swap(cp); T.chain_first.swap(cf); T.chain_len.swap(cl);
}

