// Invariant check of pygsti_b200/csrc/trie_host.h (host-side trie + heavy-path chains) on random circuit sets:
// every node covered by exactly one chain, a chain's parent chain is handed out earlier (the deadlock-freedom condition
// of k_trie_chains), every circuit's depth->node path is consistent with its symbols, node count == number of distinct
// prefixes.  Built and run by tests/test_trie_host.py with g++ (no GPU).
#include <cstdio>
#include <map>
#include <random>
#include "../pygsti_b200/csrc/trie_host.h"

int main(){
  std::mt19937 rng(1);
  for(int trial=0;trial<200;++trial){
    int n=1+rng()%300; int nsym=1+rng()%4; int nroot=1+rng()%3; bool rev=trial&1;
    std::vector<int32_t> root(n), sym; std::vector<uint32_t> ptr(n+1,0);
    for(int c=0;c<n;++c){ root[c]=rev?0:rng()%nroot; int L=rng()%12; if(trial%7==0) L=rng()%3; for(int k=0;k<L;++k) sym.push_back(rng()%nsym); ptr[c+1]=sym.size(); }
    TrieHost T; build_trie(n,root,ptr,sym,rev,T);
    size_t N=T.node_op.size();
    // parent map
    std::vector<int64_t> par(N,-100); std::vector<int> chain_of(N,-1);
    std::vector<int> seen(N,0);
    for(size_t k=0;k<T.chain_first.size();++k){
      for(uint32_t i=0;i<T.chain_len[k];++i){ uint32_t id=T.chain_first[k]+i; if(id>=N||seen[id]){printf("FAIL overlap\n");return 1;} seen[id]=1; chain_of[id]=k; par[id]= i? (int64_t)id-1 : (int64_t)T.chain_parent[k]; }
      if(T.chain_len[k]==0){printf("FAIL empty chain\n");return 1;}
      if(T.chain_parent[k]>=0){ int pc=chain_of[T.chain_parent[k]]; if(pc<0||pc>=(int)k){printf("FAIL order: parent chain %d of chain %zu\n",pc,k);return 1;} }
    }
    for(size_t i=0;i<N;++i) if(!seen[i]){printf("FAIL uncovered\n");return 1;}
    for(int c=0;c<n;++c){ uint32_t L=ptr[c+1]-ptr[c];
      uint32_t r=T.depth_node[T.dptr[c]]; if(par[r]!=-(1+root[c])||T.node_op[r]!=255){printf("FAIL root\n");return 1;}
      for(uint32_t d=1;d<=L;++d){ uint32_t id=T.depth_node[T.dptr[c]+d]; int32_t sy= rev? sym[ptr[c]+(L-d)] : sym[ptr[c]+d-1];
        if(par[id]!=(int64_t)T.depth_node[T.dptr[c]+d-1]||T.node_op[id]!=(uint8_t)sy){printf("FAIL path c=%d d=%u\n",c,d);return 1;} } }
    // sharing: number of nodes == number of distinct (root,prefix)
    std::map<std::vector<int>,int> pf;
    for(int c=0;c<n;++c){ uint32_t L=ptr[c+1]-ptr[c]; std::vector<int> key{root[c]}; pf[key]=1; for(uint32_t d=0;d<L;++d){ key.push_back(rev? sym[ptr[c]+(L-1-d)]:sym[ptr[c]+d]); pf[key]=1; } }
    if(pf.size()!=N){printf("FAIL node count %zu vs %zu\n",pf.size(),N);return 1;}
  }
  printf("build_trie OK\n"); return 0; }
