// BASELINE config 4 played call by call from compiled code — the same closed loop as
// hivedscheduler_b200/trace.py::run_c4_interactive (the harness plays kube-scheduler —
// Filtering-phase Schedule; on a preempt result Preempting-phase Schedule, delete every pod of every victim gang,
// Filtering-phase Schedule again), without an interpreter between the calls.  Test / bench infrastructure: it drives
// ANY implementation of include/hived.h through the hived_process_events pointer it is given (product or oracle),
// and returns the decision log so that the Python side can hash it exactly like the Python harness's log.
//
// Events that do not depend on each other's answers travel together (the deletions of one gang's pods: at most
// HIVED_MAX_MEMBERS per call): the ABI defines a batch as the calls one by one, in order.
#include <cstdint>
#include <cstring>
#include <chrono>
#include <deque>
#include <set>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include "hived.h"

namespace {
struct Rng {  // xorshift64* of SURVEY.md section 8d (trace.py::XorShift64Star)
  uint64_t x;
  explicit Rng(uint64_t seed) : x(seed ? seed : 0x9E3779B97F4A7C15ull) {}
  uint64_t next() { x ^= x >> 12; x ^= x << 25; x ^= x >> 27; return x * 0x2545F4914F6CDD1Dull; }
  int below(int n) { return (int)(next() % (uint64_t)n); }
};
void gangShape(int r, int& podNum, int& leafNum) {  // trace.py::_gang_shape
  if (r < 40) { podNum = 1; leafNum = 1; }
  else if (r < 65) { podNum = 1; leafNum = 4; }
  else if (r < 90) { podNum = 1; leafNum = 8; }
  else { podNum = 8; leafNum = 8; }
}
struct PodHome { int group, leafNum, podIndex, vc; };
}  // namespace

extern "C" {
typedef int (*hived_process_fn)(hived_ctx*, const hived_event_t*, int32_t, const uint32_t*, int64_t, hived_result_t*,
                                int32_t*, int64_t);

// log: int32 words, one record per decision: [gang, pod, kind, n, payload[n]] with kind 0 = preempt (payload: victim
// pod ids ascending), 1 = bind (payload: node, then this pod's leaf indices), 2 = wait (payload: wait code),
// 3 = preempt again (payload: wait code).  Returns 0, or the failing call's return code; -1: log capacity.
int c4_play(hived_process_fn process, hived_ctx* ctx, int32_t n_gangs, int32_t n_vcs, int32_t vc_gpus, int32_t total_gpus,
            double load, int32_t batch_deletes, int32_t* log, int64_t log_cap, int64_t* log_words, int64_t* n_calls,
            int64_t* n_events, double* seconds) {
  Rng rng(0x9E3779B97F4A7C15ull ^ 4ull);
  std::unordered_map<int, PodHome> podHome;
  std::unordered_map<int, std::vector<int>> groupPods;
  std::vector<int> groupSize(n_gangs, -1);
  std::vector<char> inAlive(n_gangs, 0);
  std::vector<std::deque<int>> alive(n_vcs);
  std::vector<long long> aliveGpus(n_vcs, 0);
  std::deque<std::pair<int, int>> oppAlive;
  long long oppGpus = 0;
  const long long oppLimit = (long long)(load * total_gpus), vcLimit = (long long)(load * vc_gpus);
  int nextPod = 0;
  int64_t lw = 0, calls = 0, events = 0;
  std::vector<hived_event_t> ev(HIVED_MAX_MEMBERS + 1);
  std::vector<hived_result_t> res(HIVED_MAX_MEMBERS + 1);
  std::vector<int32_t> pool(4096);
  int rc = 0;

  auto run = [&](int n) {
    calls++; events += n;
    return process(ctx, ev.data(), n, nullptr, 0, res.data(), pool.data(), (int64_t)pool.size());
  };
  auto fillDelete = [&](hived_event_t& e, int g, int leafNum, int podIndex, int vc) {
    std::memset(&e, 0, sizeof e);
    e.type = HIVED_EV_DELETE_ALLOCATED;
    e.arg0 = podIndex;
    e.suggested_off = -1;
    e.spec.group = g; e.spec.leaf_num = leafNum; e.spec.vc = vc;
  };
  auto deleteGroup = [&](int g, int vc) -> int {
    auto it = groupPods.find(g);
    if (it == groupPods.end()) return 0;
    std::vector<int> pods = std::move(it->second);
    groupPods.erase(it);
    int k = 0;
    for (int pid : pods) {
      const PodHome h = podHome[pid];
      podHome.erase(pid);
      fillDelete(ev[k++], g, h.leafNum, h.podIndex, vc);
      if (!batch_deletes || k == HIVED_MAX_MEMBERS) { int r = run(k); if (r) return r; k = 0; }
    }
    if (k) { int r = run(k); if (r) return r; }
    return 0;
  };
  auto logPut = [&](int g, int j, int kind, const int32_t* payload, int n) -> bool {
    if (lw + 4 + n > log_cap) return false;
    log[lw++] = g; log[lw++] = j; log[lw++] = kind; log[lw++] = n;
    for (int i = 0; i < n; i++) log[lw++] = payload[i];
    return true;
  };

  const auto t0 = std::chrono::steady_clock::now();
  for (int g = 0; g < n_gangs && rc == 0; g++) {
    int podNum, leafNum;
    gangShape(rng.below(100), podNum, leafNum);
    const int v = rng.below(n_vcs);
    const bool opportunistic = rng.below(2) == 1;
    const int prio = opportunistic ? -1 : rng.below(3);
    const int size = podNum * leafNum;
    if (opportunistic) {
      while (oppGpus + size > oppLimit && !oppAlive.empty()) {
        auto [og, ov] = oppAlive.front(); oppAlive.pop_front();
        if (groupPods.count(og)) { oppGpus -= groupSize[og]; if ((rc = deleteGroup(og, ov))) break; }
      }
    } else {
      while (aliveGpus[v] + size > vcLimit && !alive[v].empty()) {
        const int og = alive[v].front(); alive[v].pop_front(); inAlive[og] = 0;
        if (groupPods.count(og)) { aliveGpus[v] -= groupSize[og]; if ((rc = deleteGroup(og, v))) break; }
      }
    }
    if (rc) break;
    bool bound = true;
    for (int j = 0; j < podNum; j++) {
      int pid = -1;
      auto sched = [&](int phase) -> int {
        hived_event_t& e = ev[0];
        std::memset(&e, 0, sizeof e);
        e.type = HIVED_EV_SCHEDULE; e.phase = phase; e.suggested_off = -1;
        pid = nextPod++;
        e.spec.pod = pid; e.spec.group = g; e.spec.vc = v; e.spec.priority = prio; e.spec.pinned = -1;
        e.spec.leaf_type = 0; e.spec.leaf_num = leafNum; e.spec.flags = HIVED_SPEC_IGNORE_SUGGESTED;
        e.spec.n_members = 1; e.spec.member_leaf_num[0] = leafNum; e.spec.member_pod_num[0] = podNum;
        return run(1);
      };
      if ((rc = sched(HIVED_PHASE_FILTERING))) break;
      if (res[0].kind == HIVED_KIND_PREEMPT) {
        if ((rc = sched(HIVED_PHASE_PREEMPTING))) break;
        std::set<int> victims;
        for (int k = 0; k < res[0].n_victims; k++) victims.insert(pool[res[0].victim_off + 2 * k]);
        std::vector<int32_t> vv(victims.begin(), victims.end());
        if (!logPut(g, j, 0, vv.data(), (int)vv.size())) { rc = -1; break; }
        std::set<int> vgs;
        for (int p : victims) { auto it = podHome.find(p); if (it != podHome.end()) vgs.insert(it->second.group); }
        for (int vg : vgs) {
          const int gvc = podHome[groupPods[vg][0]].vc;
          if (groupSize[vg] >= 0) {
            if (gvc >= 0 && inAlive[vg]) aliveGpus[gvc] -= groupSize[vg];
            else if (gvc < 0) oppGpus -= groupSize[vg];
          }
          if ((rc = deleteGroup(vg, gvc >= 0 ? gvc : 0))) break;
        }
        if (rc) break;
        if ((rc = sched(HIVED_PHASE_FILTERING))) break;
      }
      const hived_result_t& r = res[0];
      if (r.kind == HIVED_KIND_BIND) {
        podHome[pid] = PodHome{g, leafNum, r.pod_index, opportunistic ? -1 : v};
        groupPods[g].push_back(pid);
        int32_t payload[1 + 64];
        payload[0] = r.node;
        const int n = r.this_n < 64 ? r.this_n : 64;
        for (int k = 0; k < n; k++) payload[1 + k] = pool[r.this_off + 3 * k + 1];
        if (!logPut(g, j, 1, payload, 1 + n)) { rc = -1; break; }
      } else {
        const int32_t code = r.wait_code;
        if (!logPut(g, j, r.kind == HIVED_KIND_WAIT ? 2 : 3, &code, 1)) { rc = -1; break; }
        bound = false;
        break;
      }
    }
    if (rc) break;
    if (bound) {
      groupSize[g] = size;
      if (opportunistic) { oppAlive.emplace_back(g, v); oppGpus += size; }
      else { alive[v].push_back(g); inAlive[g] = 1; aliveGpus[v] += size; }
    } else if (groupPods.count(g)) {
      rc = deleteGroup(g, v);
    }
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (log_words) *log_words = lw;
  if (n_calls) *n_calls = calls;
  if (n_events) *n_events = events;
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return rc;
}
}  // extern "C"
