// The extender's pod state machine with batch draining (include/hived_frontend.h).
// Host code above the public ABI of hived.h (hived_process_events / hived_schedule / hived_add_allocated_pod):
// it restates pkg/scheduler/scheduler.go:252-383, 423-469, 485-583, 640-721 with one change of shape — concurrent
// callers queue and ONE of them drains the queue into one ordered batch (the header has the equivalence argument).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/hived_frontend.h"

struct hived_fe {
  hived_ctx* ctx = nullptr;
  hived_ingest* ing = nullptr;
  hived_fe_config_t cfg{};
  std::mutex mu;     // queue, pods, answers, counters
  std::mutex ingMu;  // the ingest object caches the previous request body: one user at a time
  std::condition_variable cv;
  bool draining = false;

  struct PodRec {
    int state = HIVED_POD_WAITING;
    hived_pod_spec_t spec{};
    int specRc = 0;            // ExtractPodSchedulingSpec's verdict, reported when the pod is scheduled (:513)
    std::string specErr, group;
    int node = -1, chain = -1, podIndex = -1, bindAttempts = 0;
    std::vector<int32_t> leafIdx;
  };
  std::unordered_map<std::string, PodRec> pods;
  std::unordered_map<std::string, int> groupRefs;  // affinity-group name -> pods that name it

  enum { REQ_FILTER = 0, REQ_DELETE = 1, REQ_PREEMPT = 2 };
  struct Req {
    int64_t ticket = 0;
    int type = REQ_FILTER;
    std::string uid;
    std::vector<uint32_t> bitmap;  // empty: every node
  };
  std::deque<Req> queue;
  std::unordered_map<int64_t, hived_fe_response_t> answers;
  int64_t nextTicket = 1;
  int64_t nAnswered = 0, nDrains = 0, nEvents = 0, maxBatch = 0, nWaits = 0, heldMs = 0, nPerCall = 0;
  std::string err;
  int words = 0;

  static void setMsg(hived_fe_response_t& r, int kind, int rc, const std::string& m) {
    r.kind = kind;
    r.error = rc;
    snprintf(r.message, sizeof r.message, "%s", m.c_str());
  }
  bool bitSet(const std::vector<uint32_t>& bm, int node) const {
    if (bm.empty()) return true;
    return node >= 0 && (bm[(size_t)node >> 5] >> (node & 31)) & 1u;
  }
  // shouldForceBind :423-469 (the node-exists half of validatePodBindInfo is the shim's: it owns the node lister)
  bool shouldForceBind(const PodRec& p, const std::vector<uint32_t>& bm) const {
    return p.bindAttempts >= cfg.force_bind_threshold || !bitSet(bm, p.node);
  }
  void fillBind(hived_fe_response_t& r, const PodRec& p, bool insisted, const std::vector<uint32_t>& bm) {
    r.kind = HIVED_FE_BIND;
    r.node = p.node;
    r.chain = p.chain;
    r.insisted = insisted ? 1 : 0;
    r.bind_attempts = p.bindAttempts;
    r.force_bind = shouldForceBind(p, bm) ? 1 : 0;
    r.n_leaves = (int32_t)std::min<size_t>(p.leafIdx.size(), HIVED_FE_MAX_LEAVES);
    for (int i = 0; i < r.n_leaves; i++) r.leaf_index[i] = p.leafIdx[i];
  }
  void dropPod(const std::string& uid, const PodRec& p) {
    hived_ingest_release(ing, 1, uid.data(), (int32_t)uid.size());
    auto it = groupRefs.find(p.group);
    if (it != groupRefs.end() && --it->second <= 0) {
      groupRefs.erase(it);
      hived_group_info_t gi{};
      // the id goes back to the pool once the algorithm has forgotten the group (hived.h "Id lifetime")
      if (p.spec.group >= 0 && (hived_get_group(ctx, p.spec.group, &gi) != 0 || (gi.state == HIVED_GROUP_NONE && !gi.referenced)))
        hived_ingest_release(ing, 0, p.group.data(), (int32_t)p.group.size());
    }
  }

  // one batch: the head of the queue, every pod at most once, no preempt request inside.  Called with `lk` held and
  // draining == true; releases the lock around the device call.
  void drainOnce(std::unique_lock<std::mutex>& lk) {
    const int cap = cfg.max_batch > 0 ? cfg.max_batch : 4096;
    if (!queue.empty() && queue.front().type == REQ_PREEMPT) { preemptOne(lk); return; }
    std::vector<Req> reqs;
    std::unordered_set<std::string> seen;
    while (!queue.empty() && (int)reqs.size() < cap) {
      Req& q = queue.front();
      if (q.type == REQ_PREEMPT || seen.count(q.uid)) break;
      seen.insert(q.uid);
      reqs.push_back(std::move(q));
      queue.pop_front();
    }
    std::vector<hived_event_t> events;
    std::vector<int> evReq;            // event -> request
    std::vector<uint32_t> suggPool;
    std::vector<std::pair<std::string, PodRec>> deleted;
    long long poolWords = 64;
    for (size_t i = 0; i < reqs.size(); i++) {
      Req& q = reqs[i];
      auto it = pods.find(q.uid);
      if (q.type == REQ_DELETE) {
        if (it == pods.end()) continue;
        PodRec& p = it->second;
        if (p.specRc == 0) {
          hived_event_t ev{};
          ev.suggested_off = -1;
          ev.spec = p.spec;
          if (p.state == HIVED_POD_BINDING || p.state == HIVED_POD_BOUND) { ev.type = HIVED_EV_DELETE_ALLOCATED; ev.arg0 = p.podIndex; }
          else ev.type = HIVED_EV_DELETE_UNALLOCATED;
          events.push_back(ev);
          evReq.push_back((int)i);
        }
        deleted.emplace_back(q.uid, std::move(p));
        pods.erase(it);
        continue;
      }
      hived_fe_response_t r{};
      r.node = -1;
      // generalScheduleAdmissionCheck :364-383
      if (it == pods.end()) {
        setMsg(r, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, "Pod does not exist, completed or has not been informed to the scheduler");
      } else if (it->second.state == HIVED_POD_BOUND) {
        setMsg(r, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, std::string("Pod has already been bound to node ") + nodeName(it->second.node));
      } else if (it->second.state == HIVED_POD_BINDING) {
        it->second.bindAttempts++;  // insist on the previous result: binding is idempotent (:494-508)
        fillBind(r, it->second, true, q.bitmap);
      } else if (it->second.specRc != 0) {
        setMsg(r, HIVED_FE_ERROR, it->second.specRc, it->second.specErr);
      } else {
        hived_event_t ev{};
        ev.type = HIVED_EV_SCHEDULE;
        ev.phase = HIVED_PHASE_FILTERING;
        ev.spec = it->second.spec;
        if (q.bitmap.empty()) ev.suggested_off = -1;
        else { ev.suggested_off = (int64_t)suggPool.size(); suggPool.insert(suggPool.end(), q.bitmap.begin(), q.bitmap.end()); }
        long long leaves = 0;
        for (int m = 0; m < ev.spec.n_members; m++) leaves += (long long)ev.spec.member_leaf_num[m] * ev.spec.member_pod_num[m];
        poolWords += 3 * leaves + 2 * HIVED_FE_MAX_VICTIMS * 64;
        events.push_back(ev);
        evReq.push_back((int)i);
        continue;
      }
      r.batch_events = 0;
      answers[q.ticket] = r;
      nAnswered++;
    }
    int rc = 0;
    std::vector<hived_result_t> res(events.size());
    std::vector<int32_t> pool((size_t)poolWords);
    if (!events.empty()) {
      lk.unlock();
      rc = hived_process_events(ctx, events.data(), (int32_t)events.size(), suggPool.empty() ? nullptr : suggPool.data(),
                                (int64_t)suggPool.size(), res.data(), pool.data(), (int64_t)pool.size());
      lk.lock();
      nDrains++;
      nEvents += (int64_t)events.size();
      if ((int64_t)events.size() > maxBatch) maxBatch = (int64_t)events.size();
    }
    int waits = 0;
    for (size_t e = 0; e < events.size(); e++) {
      Req& q = reqs[(size_t)evReq[e]];
      if (q.type == REQ_DELETE) continue;
      hived_fe_response_t r{};
      r.node = -1;
      r.batch_events = (int32_t)events.size();
      const hived_result_t& x = res[e];
      auto it = pods.find(q.uid);
      const int evRc = rc != 0 ? rc : x.error;
      if (evRc != 0 || it == pods.end()) {
        const char* m = hived_last_error(ctx);
        setMsg(r, HIVED_FE_ERROR, evRc ? evRc : HIVED_ERR_PLATFORM, m ? m : "");
      } else if (x.kind == HIVED_KIND_BIND) {
        PodRec& p = it->second;  // :516-540: the pod is assumed allocated, Binding from now on
        p.state = HIVED_POD_BINDING;
        p.node = x.node;
        p.chain = x.chain;
        p.podIndex = x.pod_index;
        p.bindAttempts = 0;
        p.leafIdx.clear();
        for (int k = 0; k < x.this_n; k++) p.leafIdx.push_back(pool[(size_t)x.this_off + 3 * (size_t)k + 1]);
        fillBind(r, p, false, q.bitmap);
      } else if (x.kind == HIVED_KIND_PREEMPT) {
        r.kind = HIVED_FE_PREEMPT;  // :541-559: state unchanged, K8s is told that preemption may help
        r.n_victims = x.n_victims < HIVED_FE_MAX_VICTIMS ? x.n_victims : HIVED_FE_MAX_VICTIMS;
        for (int k = 0; k < r.n_victims; k++) { r.victim_pod[k] = pool[(size_t)x.victim_off + 2 * (size_t)k]; r.victim_node[k] = pool[(size_t)x.victim_off + 2 * (size_t)k + 1]; }
      } else {
        it->second.state = HIVED_POD_WAITING;  // :560-582
        r.kind = HIVED_FE_WAIT;
        r.wait_code = x.wait_code;
        r.wait_cell = x.wait_cell;
        waits++;
      }
      answers[q.ticket] = r;
      nAnswered++;
    }
    for (auto& d : deleted) dropPod(d.first, d.second);
    nWaits += waits;
    if (cfg.waiting_block_ms > 0 && waits > 0) {
      // WaitingPodSchedulingBlockMilliSec :567-571 — the scheduler is held (draining stays true), callers keep queueing
      const long long ms = (long long)cfg.waiting_block_ms * waits;
      cv.notify_all();
      lk.unlock();
      std::this_thread::sleep_for(std::chrono::milliseconds(ms));
      lk.lock();
      heldMs += ms;
    }
  }
  const char* nodeName(int node) const {
    const char* s = node >= 0 ? hived_node_name(ctx, node) : nullptr;
    return s ? s : "?";
  }
  // preemptRoutine :640-721 — Schedule in the preempting phase, no AddAllocatedPod: one call (hived_schedule)
  void preemptOne(std::unique_lock<std::mutex>& lk) {
    Req q = std::move(queue.front());
    queue.pop_front();
    hived_fe_response_t r{};
    r.node = -1;
    auto it = pods.find(q.uid);
    if (it == pods.end()) setMsg(r, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, "Pod does not exist, completed or has not been informed to the scheduler");
    else if (it->second.state == HIVED_POD_BOUND) setMsg(r, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, std::string("Pod has already been bound to node ") + nodeName(it->second.node));
    else if (it->second.state == HIVED_POD_BINDING) setMsg(r, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, std::string("Pod has already been binding to node ") + nodeName(it->second.node));
    else if (it->second.specRc != 0) setMsg(r, HIVED_FE_ERROR, it->second.specRc, it->second.specErr);
    else {
      hived_pod_spec_t sp = it->second.spec;
      hived_result_t x{};
      std::vector<int32_t> pool((size_t)(3 * 64 * 64 + 2 * 4096 + 64));
      lk.unlock();
      int rc = hived_schedule(ctx, &sp, q.bitmap.empty() ? nullptr : q.bitmap.data(), HIVED_PHASE_PREEMPTING, &x, pool.data(), (int32_t)pool.size());
      lk.lock();
      nPerCall++;
      it = pods.find(q.uid);
      if (rc != 0 || it == pods.end()) { const char* m = hived_last_error(ctx); setMsg(r, HIVED_FE_ERROR, rc ? rc : HIVED_ERR_PLATFORM, m ? m : ""); }
      else if (x.kind == HIVED_KIND_BIND) r.kind = HIVED_FE_NONE;  // free resource appeared: filterRoutine will bind it
      else if (x.kind == HIVED_KIND_PREEMPT) {
        it->second.state = HIVED_POD_PREEMPTING;
        r.kind = HIVED_FE_PREEMPT;
        r.n_victims = x.n_victims < HIVED_FE_MAX_VICTIMS ? x.n_victims : HIVED_FE_MAX_VICTIMS;
        for (int k = 0; k < r.n_victims; k++) { r.victim_pod[k] = pool[(size_t)x.victim_off + 2 * (size_t)k]; r.victim_node[k] = pool[(size_t)x.victim_off + 2 * (size_t)k + 1]; }
      } else {
        it->second.state = HIVED_POD_WAITING;
        r.kind = HIVED_FE_NONE;
        r.wait_code = x.wait_code;
        r.wait_cell = x.wait_cell;
      }
    }
    answers[q.ticket] = r;
    nAnswered++;
  }

  int64_t enqueue(int type, const char* uid, const char* json, int64_t len) {
    Req q;
    q.type = type;
    q.uid = uid ? uid : "";
    if (json && len > 0) {
      std::lock_guard<std::mutex> g(ingMu);
      q.bitmap.assign((size_t)words, 0u);
      int32_t all = 0;
      int32_t cnt = hived_ingest_node_names_json(ing, json, len, q.bitmap.data(), &all, nullptr, nullptr);
      if (cnt < 0) { std::lock_guard<std::mutex> g2(mu); err = hived_ingest_last_error(ing); return -1; }
      if (all) q.bitmap.clear();  // every node of the cluster: no bitmap at all
    }
    std::lock_guard<std::mutex> g(mu);
    q.ticket = nextTicket++;
    const int64_t t = q.ticket;
    queue.push_back(std::move(q));
    return t;
  }
  // leader / follower: whoever finds the scheduler idle drains; the others sleep until their answer is there
  int await(int64_t ticket, hived_fe_response_t* out) {
    std::unique_lock<std::mutex> lk(mu);
    while (true) {
      auto it = answers.find(ticket);
      if (it != answers.end()) { *out = it->second; answers.erase(it); return 0; }
      if (!draining) {
        draining = true;
        drainOnce(lk);
        draining = false;
        cv.notify_all();
        continue;
      }
      cv.wait(lk);
    }
  }
};

extern "C" {

int hived_fe_create(hived_ctx* ctx, hived_ingest* ing, const hived_fe_config_t* cfg, hived_fe** out) {
  *out = nullptr;
  if (!ctx || !ing || !cfg) return HIVED_ERR_BAD_SPEC;
  hived_fe* f = new hived_fe();
  f->ctx = ctx;
  f->ing = ing;
  f->cfg = *cfg;
  if (f->cfg.force_bind_threshold <= 0) f->cfg.force_bind_threshold = 3;
  f->words = hived_ingest_bitmap_words(ing);
  *out = f;
  return 0;
}
void hived_fe_destroy(hived_fe* f) { delete f; }
const char* hived_fe_last_error(hived_fe* f) { return f->err.c_str(); }

// the annotation's gang name defaults to "namespace/name" (internal/utils.go:249-256) while ids are keyed by UID:
// parse with the key as the default gang name, then intern the pod under its UID
static int fe_parse(hived_fe* f, const char* uid, const char* key, const char* ann, int64_t annLen, hived_fe::PodRec& p) {
  std::lock_guard<std::mutex> g(f->ingMu);
  const char* k = (key && *key) ? key : uid;
  p.specRc = hived_ingest_pod_spec_yaml(f->ing, ann, annLen, k, f->cfg.max_groups, f->cfg.max_pods, &p.spec);
  if (p.specRc != 0) { p.specErr = hived_ingest_last_error(f->ing); return p.specRc; }
  // pod id: by UID (the helper interned the key; swap when they differ)
  if (strcmp(k, uid) != 0) {
    hived_ingest_release(f->ing, 1, k, -1);
    p.spec.pod = hived_ingest_intern(f->ing, 1, uid, -1, f->cfg.max_pods);
    if (p.spec.pod < 0) { p.specRc = HIVED_ERR_CAPACITY; p.specErr = "pod id table full"; return p.specRc; }
  }
  p.group = hived_ingest_last_group_name(f->ing);  // for the reference count that recycles the gang's id
  return 0;
}

int hived_fe_add_unbound_pod(hived_fe* f, const char* uid, const char* key, const char* ann, int64_t annLen) {
  if (!uid) return HIVED_ERR_BAD_SPEC;
  {
    std::lock_guard<std::mutex> g(f->mu);
    if (f->pods.count(uid)) return 0;  // keep the existing one (:343-347)
  }
  hived_fe::PodRec p;
  p.state = HIVED_POD_WAITING;
  fe_parse(f, uid, key, ann, annLen, p);
  std::lock_guard<std::mutex> g(f->mu);
  if (f->pods.count(uid)) return 0;
  if (p.specRc == 0) f->groupRefs[p.group]++;
  f->pods.emplace(uid, std::move(p));
  return 0;
}

int hived_fe_add_bound_pod(hived_fe* f, const char* uid, const char* key, const char* ann, int64_t annLen, const hived_bind_info_t* info,
                           const int32_t* leaves, int32_t nLeafInts) {
  if (!uid) return HIVED_ERR_BAD_SPEC;
  {
    std::lock_guard<std::mutex> g(f->mu);
    auto it = f->pods.find(uid);
    if (it != f->pods.end() && (it->second.state == HIVED_POD_BINDING || it->second.state == HIVED_POD_BOUND)) {
      it->second.state = HIVED_POD_BOUND;  // already allocated: the placement never changes (:316-326)
      return 0;
    }
  }
  if (!info) { std::lock_guard<std::mutex> g(f->mu); f->err = "recovering a bound pod needs its PodBindInfo"; return HIVED_ERR_BAD_SPEC; }
  hived_fe::PodRec p;
  if (fe_parse(f, uid, key, ann, annLen, p) != 0) { std::lock_guard<std::mutex> g(f->mu); f->err = p.specErr; return p.specRc; }
  // recover (:329-335).  The scheduler is held like for any other mutation.
  std::unique_lock<std::mutex> lk(f->mu);
  while (f->draining) f->cv.wait(lk);
  f->draining = true;
  lk.unlock();
  int rc = hived_add_allocated_pod(f->ctx, &p.spec, info, leaves, nLeafInts);
  lk.lock();
  f->draining = false;
  f->cv.notify_all();
  if (rc != 0) { const char* m = hived_last_error(f->ctx); f->err = m ? m : ""; return rc; }
  p.state = HIVED_POD_BOUND;
  p.node = info->node;
  p.chain = info->chain;
  // getAllocatedPodIndex (utils.go:291-304): the row of this pod among those with its leaf number
  hived_group_placement_t gp{};
  std::vector<int32_t> gl((size_t)64 * 64), gv((size_t)64 * 64), pods(256), pre(256);
  p.podIndex = 0;
  if (hived_get_group_placement(f->ctx, p.spec.group, &gp, gl.data(), gv.data(), (int32_t)gl.size(), pods.data(), (int32_t)pods.size(),
                                pre.data(), (int32_t)pre.size()) == 0) {
    int row = 0;
    for (int m = 0; m < gp.n_members; m++) {
      for (int j = 0; j < gp.member_pod_num[m]; j++, row++)
        if (gp.member_leaf_num[m] == p.spec.leaf_num && row < gp.n_pods && pods[(size_t)row] == p.spec.pod) p.podIndex = j;
    }
  }
  f->groupRefs[p.group]++;
  f->pods[uid] = std::move(p);
  return 0;
}

int hived_fe_delete_pod(hived_fe* f, const char* uid) {
  if (!uid) return HIVED_ERR_BAD_SPEC;
  hived_fe::Req q;
  q.type = hived_fe::REQ_DELETE;
  q.uid = uid;
  std::lock_guard<std::mutex> g(f->mu);
  q.ticket = 0;
  f->queue.push_back(std::move(q));
  return 0;
}

int32_t hived_fe_pod_state(hived_fe* f, const char* uid) {
  std::lock_guard<std::mutex> g(f->mu);
  auto it = f->pods.find(uid ? uid : "");
  return it == f->pods.end() ? HIVED_POD_UNKNOWN : it->second.state;
}

int hived_fe_filter(hived_fe* f, const char* uid, const char* json, int64_t len, hived_fe_response_t* out) {
  int64_t t = f->enqueue(hived_fe::REQ_FILTER, uid, json, len);
  if (t < 0) { memset(out, 0, sizeof *out); hived_fe::setMsg(*out, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, f->err); return HIVED_ERR_BAD_SPEC; }
  return f->await(t, out);
}
int hived_fe_preempt(hived_fe* f, const char* uid, const char* json, int64_t len, hived_fe_response_t* out) {
  int64_t t = f->enqueue(hived_fe::REQ_PREEMPT, uid, json, len);
  if (t < 0) { memset(out, 0, sizeof *out); hived_fe::setMsg(*out, HIVED_FE_ERROR, HIVED_ERR_BAD_SPEC, f->err); return HIVED_ERR_BAD_SPEC; }
  return f->await(t, out);
}
int64_t hived_fe_enqueue_filter(hived_fe* f, const char* uid, const char* json, int64_t len) {
  return f->enqueue(hived_fe::REQ_FILTER, uid, json, len);
}
int hived_fe_drain(hived_fe* f) {
  std::unique_lock<std::mutex> lk(f->mu);
  while (f->draining) f->cv.wait(lk);
  f->draining = true;
  while (!f->queue.empty()) f->drainOnce(lk);
  f->draining = false;
  f->cv.notify_all();
  return 0;
}
int hived_fe_take(hived_fe* f, int64_t ticket, hived_fe_response_t* out) {
  std::lock_guard<std::mutex> g(f->mu);
  auto it = f->answers.find(ticket);
  if (it == f->answers.end()) return HIVED_ERR_BAD_SPEC;
  *out = it->second;
  f->answers.erase(it);
  return 0;
}
int hived_fe_bind_check(hived_fe* f, const char* uid, int32_t node, char* message, int32_t cap) {
  std::lock_guard<std::mutex> g(f->mu);
  auto it = f->pods.find(uid ? uid : "");
  std::string m;
  if (it == f->pods.end()) m = "Pod does not exist, completed or has not been informed to the scheduler";
  else if (it->second.state == HIVED_POD_BOUND) m = std::string("Pod has already been bound to node ") + f->nodeName(it->second.node);
  else if (it->second.state == HIVED_POD_BINDING) {
    if (it->second.node == node) return 0;
    m = std::string("Pod binding node mismatch: expected ") + f->nodeName(it->second.node) + ", received " + f->nodeName(node);
  } else m = "Pod cannot be bound without a scheduling placement";
  if (message && cap > 0) snprintf(message, (size_t)cap, "%s", m.c_str());
  return HIVED_ERR_BAD_SPEC;
}
int hived_fe_stats(hived_fe* f, int64_t* out, int32_t n) {
  std::lock_guard<std::mutex> g(f->mu);
  const int64_t v[7] = {f->nAnswered, f->nDrains, f->nEvents, f->maxBatch, f->nWaits, f->heldMs, f->nPerCall};
  for (int i = 0; i < n && i < 7; i++) out[i] = v[i];
  return 7;
}

}  // extern "C"
