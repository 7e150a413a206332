/* Per-call latency of the extender-shaped path, from plain C (no Python in the timed loop):
 *   filterRoutine's sequence for ONE pod = hived_process_events(n = 1, HIVED_EV_SCHEDULE)   (Schedule + AddAllocatedPod)
 *   deletePod                             = hived_delete_allocated_pod
 * on the 64k-GPU cluster of BASELINE configs[2] with the VCs half full (so that the cluster views have work to do).
 *
 *   gcc -O2 -I include -o percall profiles/micro/percall_latency.c -L hivedscheduler_b200/csrc -lhived_cuda -Wl,-rpath,...
 *   ./percall c3.spec [calls]          HIVED_NO_RESIDENT=1 ./percall c3.spec      (one kernel launch per call)
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hived.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}
static int cmp(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return x < y ? -1 : x > y; }
static void report(const char* what, double* v, int n) {
  qsort(v, (size_t)n, sizeof(double), cmp);
  double sum = 0;
  for (int i = 0; i < n; i++) sum += v[i];
  printf("{\"call\": \"%s\", \"n\": %d, \"min_us\": %.2f, \"median_us\": %.2f, \"mean_us\": %.2f, \"p99_us\": %.2f}\n", what, n, v[0], v[n / 2],
         sum / n, v[(int)(n * 0.99)]);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s spec-file [calls]\n", argv[0]); return 2; }
  const int calls = argc > 2 ? atoi(argv[2]) : 5000;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  char* spec = (char*)malloc((size_t)len + 1);
  if (fread(spec, 1, (size_t)len, f) != (size_t)len) return 2;
  spec[len] = 0;
  fclose(f);
  hived_options_t opt;
  memset(&opt, 0, sizeof opt);
  opt.max_groups = 1 << 16; opt.max_pods = 1 << 18; opt.max_group_leaves = 64; opt.max_group_pods = 8;
  hived_ctx* ctx = NULL;
  int rc = hived_create(spec, &opt, &ctx);
  if (rc) { fprintf(stderr, "hived_create: %d %s\n", rc, hived_create_error()); return 1; }
  const int nodes = hived_num_nodes(ctx), vcs = hived_num_vcs(ctx);
  /* every node healthy, then fill: 300 8-GPU pods per VC (one batch) */
  const int fill = 300 * vcs, nev = nodes + fill;
  hived_event_t* ev = (hived_event_t*)calloc((size_t)nev, sizeof *ev);
  hived_result_t* res = (hived_result_t*)calloc((size_t)nev, sizeof *res);
  const long long cap = 3ll * 8 * nev + 4096;
  int32_t* pool = (int32_t*)calloc((size_t)cap, 4);
  for (int i = 0; i < nodes; i++) { ev[i].type = HIVED_EV_NODE_HEALTH; ev[i].arg0 = i; ev[i].arg1 = 1; ev[i].suggested_off = -1; }
  for (int i = 0; i < fill; i++) {
    hived_event_t* e = &ev[nodes + i];
    e->type = HIVED_EV_SCHEDULE; e->phase = HIVED_PHASE_FILTERING; e->suggested_off = -1;
    e->spec.pod = i; e->spec.group = i; e->spec.vc = i % vcs; e->spec.priority = 0; e->spec.pinned = -1; e->spec.leaf_type = 0;
    e->spec.leaf_num = 8; e->spec.flags = HIVED_SPEC_IGNORE_SUGGESTED; e->spec.n_members = 1;
    e->spec.member_leaf_num[0] = 8; e->spec.member_pod_num[0] = 1;
  }
  rc = hived_process_events(ctx, ev, nev, NULL, 0, res, pool, cap);
  if (rc) { fprintf(stderr, "fill: %d %s\n", rc, hived_last_error(ctx)); return 1; }
  double* ts = (double*)malloc(sizeof(double) * (size_t)calls);
  double* td = (double*)malloc(sizeof(double) * (size_t)calls);
  double* to = (double*)malloc(sizeof(double) * (size_t)calls);
  hived_event_t one;
  hived_result_t r1;
  for (int k = -200; k < calls; k++) { /* 200 warm-up rounds */
    const int id = fill + (k + 200) % 1000;
    memset(&one, 0, sizeof one);
    one.type = HIVED_EV_SCHEDULE; one.phase = HIVED_PHASE_FILTERING; one.suggested_off = -1;
    one.spec.pod = id; one.spec.group = id; one.spec.vc = (k + 200) % vcs; one.spec.priority = 0; one.spec.pinned = -1; one.spec.leaf_type = 0;
    one.spec.leaf_num = 8; one.spec.flags = HIVED_SPEC_IGNORE_SUGGESTED; one.spec.n_members = 1;
    one.spec.member_leaf_num[0] = 8; one.spec.member_pod_num[0] = 1;
    /* Schedule only (what /v1/extender/preempt and a dry run cost) */
    double t0 = now_us();
    rc = hived_schedule(ctx, &one.spec, NULL, HIVED_PHASE_FILTERING, &r1, pool, (int32_t)cap);
    double t1 = now_us();
    if (rc || r1.kind != HIVED_KIND_BIND) { fprintf(stderr, "schedule: rc %d kind %d\n", rc, r1.kind); return 1; }
    rc = hived_process_events(ctx, &one, 1, NULL, 0, &r1, pool, cap);
    double t2 = now_us();
    if (rc || r1.kind != HIVED_KIND_BIND) { fprintf(stderr, "filter: rc %d kind %d %s\n", rc, r1.kind, hived_last_error(ctx)); return 1; }
    rc = hived_delete_allocated_pod(ctx, id, 8, r1.pod_index);
    double t3 = now_us();
    if (rc) { fprintf(stderr, "delete: %d\n", rc); return 1; }
    if (k >= 0) { to[k] = t1 - t0; ts[k] = t2 - t1; td[k] = t3 - t2; }
  }
  printf("{\"backend\": \"%s\", \"nodes\": %d, \"vcs\": %d, \"resident\": %s}\n", hived_backend(), nodes, vcs,
         getenv("HIVED_NO_RESIDENT") ? "false" : "true");
  report("hived_schedule (Schedule only)", to, calls);
  report("hived_process_events n=1 (Schedule + AddAllocatedPod)", ts, calls);
  report("hived_delete_allocated_pod", td, calls);
  hived_destroy(ctx);
  return 0;
}
