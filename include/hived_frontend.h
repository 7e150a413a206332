/* hived_frontend.h — the extender's pod state machine with batch draining.
 *
 * What it replaces in the reference: pkg/scheduler/scheduler.go
 *   :252-363  addPod / updatePod / deletePod / addBoundPod / addUnboundPod      (podScheduleStatuses)
 *   :364-383  generalScheduleAdmissionCheck
 *   :423-469  shouldForceBind (+ validatePodBindInfo :385-421: the bind node must be among the suggested nodes)
 *   :485-583  filterRoutine, including WaitingPodSchedulingBlockMilliSec (:567-571)
 *   :640-721  preemptRoutine
 * The reference answers one filter call at a time under schedulerLock (:486).  Here concurrent callers QUEUE their
 * requests and one of them drains the queue: every queued filter call and pod deletion becomes one event of ONE
 * hived_process_events batch (HIVED_EV_SCHEDULE is exactly filterRoutine's Schedule + AddAllocatedPod), in arrival
 * order — so the answers are those of the reference serving the same arrival order one by one, while the GPU sees
 * batches instead of single events (group commit).  A batch holds every pod at most once: a second request for a
 * pod already in the batch (its admission check depends on the first one's result) starts the next batch.
 *
 * WaitingPodSchedulingBlockMilliSec: the reference sleeps that long, holding the scheduler lock, after every
 * decision that made a pod wait ("block the whole scheduling to achieve better FIFO"): nothing is scheduled or
 * deleted meanwhile, requests queue on the lock.  A sleep changes no state, so the answers of requests that were
 * already queued do not depend on it; what it changes is WHEN later arrivals are served.  The drainer therefore
 * holds the scheduler for waiting_block_ms x (number of WAIT answers of the batch) after answering the batch:
 * the same total stall, the same arrival-order answers.
 *
 * Thread-safe: hived_fe_filter / hived_fe_preempt / the informer calls may come from any number of threads (the
 * HTTP handlers and the informers of the shim).  The Kubernetes client is the shim's: a response with force_bind=1
 * asks the caller to run its bind executor (forceBindExecutor, :471-483).                                            */
#ifndef HIVED_FRONTEND_H_
#define HIVED_FRONTEND_H_
#include "hived.h"
#include "hived_ingest.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct hived_fe hived_fe;

typedef struct hived_fe_config {
  int32_t waiting_block_ms;      /* WaitingPodSchedulingBlockMilliSec (api/config.go: default 0) */
  int32_t force_bind_threshold;  /* ForcePodBindThreshold (default 3) */
  int32_t max_batch;             /* events per drain (<= 0: 4096) */
  int32_t max_groups, max_pods;  /* the id capacities the scheduler was created with (hived_options_t) */
  int32_t reserved[3];
} hived_fe_config_t;

/* internal.PodState (pkg/internal/types.go:154-194) */
#define HIVED_POD_UNKNOWN 0
#define HIVED_POD_WAITING 1
#define HIVED_POD_PREEMPTING 2
#define HIVED_POD_BINDING 3
#define HIVED_POD_BOUND 4

#define HIVED_FE_MAX_LEAVES 64
#define HIVED_FE_MAX_VICTIMS 64
#define HIVED_FE_BIND 1      /* ExtenderFilterResult.NodeNames = [node]: a new placement, or the previous one insisted on */
#define HIVED_FE_WAIT 2      /* FailedNodes[hivedscheduler] = wait reason */
#define HIVED_FE_PREEMPT 3   /* filter: FailedNodes per victim node; preempt: NodeNameToMetaVictims */
#define HIVED_FE_ERROR 4     /* the routine panicked: error = HIVED_ERR_* (1..99: HTTP 400, >= 100: 500), message says why */
#define HIVED_FE_NONE 5      /* preempt: empty ExtenderPreemptionResult (free resource appeared, or still waiting) */
typedef struct hived_fe_response {
  int32_t kind;
  int32_t error;
  int32_t node;            /* BIND: node id to bind to */
  int32_t insisted;        /* BIND: 1 = the pod was already Binding, the previous decision is repeated */
  int32_t force_bind;      /* BIND: shouldForceBind said yes */
  int32_t bind_attempts;
  int32_t chain;           /* BIND: PodBindInfo.CellChain */
  int32_t n_leaves;        /* BIND: this pod's LeafCellIsolation */
  int32_t leaf_index[HIVED_FE_MAX_LEAVES];
  int32_t wait_code, wait_cell;  /* WAIT: HIVED_WAIT_* | scope, cell named by the reason */
  int32_t n_victims;       /* PREEMPT: (pod id, node id) pairs, all victims on all nodes (hived_result_t) */
  int32_t victim_pod[HIVED_FE_MAX_VICTIMS];
  int32_t victim_node[HIVED_FE_MAX_VICTIMS];
  int32_t batch_events;    /* how many events the drain that answered this request carried (observability) */
  char message[160];
} hived_fe_response_t;

int hived_fe_create(hived_ctx*, hived_ingest*, const hived_fe_config_t*, hived_fe** out);
void hived_fe_destroy(hived_fe*);

/* informer side (:252-363).  uid: pod UID; key: "namespace/name" (the default gang's name); annotation: the
 * pod-scheduling-spec YAML.  add_bound_pod: a pod the scheduler already allocated becomes Bound; an unknown one is
 * RECOVERED with the PodBindInfo the shim parsed from the pod-bind-info annotation (hived_add_allocated_pod).      */
int hived_fe_add_unbound_pod(hived_fe*, const char* uid, const char* key, const char* annotation, int64_t annotation_len);
int hived_fe_add_bound_pod(hived_fe*, const char* uid, const char* key, const char* annotation, int64_t annotation_len,
                           const hived_bind_info_t* info, const int32_t* leaves, int32_t n_leaf_ints);
int hived_fe_delete_pod(hived_fe*, const char* uid);   /* queued: ordered with the filter calls */
int32_t hived_fe_pod_state(hived_fe*, const char* uid);

/* extender side.  node_names_json: the NodeNames array of the request body (hived_ingest_json_find + the raw body),
 * or NULL = every node.  Blocks until a drain has answered the request (the caller may itself become the drainer). */
int hived_fe_filter(hived_fe*, const char* uid, const char* node_names_json, int64_t len, hived_fe_response_t* out);
int hived_fe_preempt(hived_fe*, const char* uid, const char* node_names_json, int64_t len, hived_fe_response_t* out);
/* bindRoutine's check (:585-617): 0 = the pod is Binding to that node, else HIVED_ERR_BAD_SPEC with the message */
int hived_fe_bind_check(hived_fe*, const char* uid, int32_t node, char* message, int32_t message_cap);

/* non-blocking pair for a caller that batches on its own: enqueue returns a ticket; drain answers everything queued
 * (at most max_batch events per hived_process_events call, as many calls as needed); take fetches and frees an answer */
int64_t hived_fe_enqueue_filter(hived_fe*, const char* uid, const char* node_names_json, int64_t len);
int hived_fe_drain(hived_fe*);
int hived_fe_take(hived_fe*, int64_t ticket, hived_fe_response_t* out);   /* 0, or HIVED_ERR_BAD_SPEC: no such answer */

/* counters: [0] requests answered, [1] drains (hived_process_events calls), [2] events in them, [3] largest batch,
 * [4] WAIT answers, [5] milliseconds the scheduler was held for waiting_block_ms, [6] per-call fallbacks (preempt) */
int hived_fe_stats(hived_fe*, int64_t* out, int32_t n);
const char* hived_fe_last_error(hived_fe*);
#ifdef __cplusplus
}
#endif
#endif
