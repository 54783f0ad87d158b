// Digest of everything mk::build_tables produces (merkurio_b200/csrc/mk_tables.h: seed geometry, filters, cuckoo
// table, postings, compare form of the patterns) for a fixed set of seeded query lists, both encodings. The
// expected lines (tests/golden/table_digests.txt) were printed by the builder as it was when the GPU parity
// suite last ran against it, before its seed index was rewritten (radix sort, rolling codes, prefetching): the
// device sees byte-identical tables, whatever the builder does to get there. Host-only code, no GPU needed.
#include "mk_tables.h"
#include <chrono>
#include <cstdio>
static double now(){return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}
static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull){ const uint8_t* b=(const uint8_t*)p; for(size_t i=0;i<n;++i){h^=b[i];h*=1099511628211ull;} return h; }
template<class T> uint64_t hv(const std::vector<T>& v){ return fnv(v.data(), v.size()*sizeof(T)) ^ v.size(); }
int main(int argc,char**argv){
  // cases: n, lo, hi, alphabet with N?, case-insensitive
  struct C{uint32_t n,lo,hi;bool withN;bool ci;} cases[]={{2000,31,31,false,false},{10000,31,31,false,false},{2000,27,27,false,false},{3000,15,18,false,false},{500,12,14,false,false},
    {200000,21,63,true,false},{60000,31,40,false,true},{50,3,9,true,false},{300000,19,25,false,false},{1000000,21,63,true,false}};
  int upto = argc>1? atoi(argv[1]) : 9;
  for(int ci=0; ci<upto; ++ci){ C c=cases[ci];
    std::mt19937_64 rng(ci+11);
    std::vector<std::string> pats(c.n);
    for(auto&p:pats){ int L=c.lo+rng()%(c.hi-c.lo+1); p.resize(L); for(auto&ch:p) ch="ACGT"[rng()&3]; if(c.withN && rng()%50==0) p[rng()%L]='N'; if(c.ci && rng()%3==0) for(auto&ch:p) ch|=0x20; }
    std::sort(pats.begin(),pats.end()); pats.erase(std::unique(pats.begin(),pats.end()),pats.end());
    std::string blob; std::vector<uint32_t> off{0};
    for(auto&p:pats){blob+=p;off.push_back(blob.size());}
    mk::PatternSet ps = mk::make_pattern_set((const uint8_t*)blob.data(), off.data(), (uint32_t)pats.size(), c.ci);
    for(int enc=0;enc<2;++enc){
      double t0=now();
      mk::Tables t = mk::build_tables(ps, enc);
      double dt=now()-t0;
      printf("case %d enc %d: q=%u d=%u q2=%u lml=%u perm=%d win=%d wm=%x/%x dual=%d seeds=%u flb=%u fh=%u fb=%u smem=%d f32=%d f2lb=%u bm=%u dperm=%d direct=%d/%u gate=%x/%x | filter %016llx filter2 %016llx slots %016llx post %016llx pb %016llx po %016llx pl %016llx",
        ci, enc, t.q,t.d,t.q2,t.long_min_len,t.perm,t.win,t.win_mask0,t.win_mask1,t.filter_dual,t.n_seeds,t.filter_log2_bits,t.filter_hashes,t.filter_blocks,t.filter_in_smem,t.filter32,t.filter2_log2_bits,t.bucket_mask,
        t.dual_perm,t.filter_direct,t.direct_q1,t.gate_mask,t.gate_val,
        (unsigned long long)hv(t.filter),(unsigned long long)hv(t.filter2),(unsigned long long)hv(t.slots),(unsigned long long)hv(t.postings),(unsigned long long)hv(t.pat_bytes),(unsigned long long)hv(t.pat_off),(unsigned long long)hv(t.pat_live));
      fprintf(stderr,"case %d enc %d: %.3f s\n",ci,enc,dt);
      printf("\n");
    }
  }
}
