// tools/plan_hash.cpp -- development check of the host plan stage: builds a plan from a synthetic workload through
// the C ABI of the library given as argv[1] (argv[2] = workloads/libmiso_synth.so, argv[3] = kind 0 / 1) and prints a
// hash of the tile arena, the descriptor offsets and the per-gene read counts.  Two builds of plan.cpp that print the
// same hash hand the kernels the same bytes.
//   g++ -std=c++17 -O1 -Imiso_b200/csrc -I/usr/local/cuda/include -o /tmp/plan_hash tools/plan_hash.cpp -ldl
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <dlfcn.h>
#include "plan.hpp"
#include "../../include/miso_synth.h"
using namespace misob200;
static uint64_t fnv(const void *p, size_t n, uint64_t h) { const unsigned char *b = (const unsigned char *) p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } return h; }
int main(int argc, char **argv) {
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  void *syn = dlopen(argv[2], RTLD_NOW | RTLD_GLOBAL);
  if (!lib || !syn) { printf("dlopen: %s\n", dlerror()); return 1; }
  auto create = (int (*)(misob200_plan_t **)) dlsym(lib, "misob200_plan_create");
  auto append = (int (*)(misob200_plan_t *, const misob200_reads_t *, int)) dlsym(lib, "misob200_plan_append");
  auto wcreate = (decltype(&misob200_workload_create)) dlsym(syn, "misob200_workload_create");
  auto wview = (decltype(&misob200_workload_view)) dlsym(syn, "misob200_workload_view");
  misob200_plan_t *plan = nullptr;
  create(&plan);
  for (int part = 0; part < 2; part++) {         // two appends: the second starts from a non-empty arena
    misob200_workload_t *w = nullptr;
    const int kind = argc > 3 ? atoi(argv[3]) : 1;
    if (wcreate(kind, 3000, kind ? 2000 : 1000, 36, 250., 900., 4., 1, 100 + 3000 * part, 8, &w)) { printf("workload failed\n"); return 1; }
    misob200_reads_t v;
    wview(w, &v);
    int rc = append(plan, &v, 8);
    if (rc) { printf("append rc %d\n", rc); return 1; }
  }
  Plan &p = plan->p;
  uint64_t h = fnv(p.tiles.data(), p.tiles.size(), 1469598103934665603ull);
  for (auto &d : p.desc) { h = fnv(&d.tile_off, sizeof(d.tile_off), h); h = fnv(&d.drawn_off, sizeof(d.drawn_off), h); h = fnv(&d.R2, sizeof(d.R2), h); }
  for (auto &g : p.host) { h = fnv(&g.read_base, sizeof(g.read_base), h); h = fnv(&g.R2, sizeof(g.R2), h); }
  printf("genes %zu tiles %zu bytes n_reads %lld n_drawn %lld hash %016llx\n", p.desc.size(), p.tiles.size(), (long long) p.n_reads, (long long) p.n_drawn, (unsigned long long) h);
  return 0;
}
