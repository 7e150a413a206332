/* hived_ingest.h — request ingest for large clusters, host-side C helpers exported by
 * libhived_cuda.so next to the ABI of hived.h.
 *
 * What they replace in the reference, per scheduling request:
 *   - pkg/algorithm/hived_algorithm.go:190-193   suggestedNodes []string -> common.Set (one map insert per name;
 *                                                 8192 names per request on the 64k-GPU cluster)
 *   - pkg/webserver/webserver.go:173-182         ExtenderArgs JSON decode (the NodeNames array)
 *   - pkg/internal/utils.go:230-242              the pod's scheduling-spec annotation (YAML) -> PodSchedulingSpec
 *
 * The wire formats do not change.  Node names are interned once per scheduler (open-addressing table over the
 * names of hived_node_name()); a request's names become the node bitmap hived_schedule() takes; a request naming
 * EVERY node of the cluster (the common case: kube-scheduler found all nodes feasible) is recognised so that the
 * caller can pass no bitmap at all (hived_schedule(suggested = NULL): every node counts as suggested,
 * topology_aware_scheduler.go:218-222); a request body identical to the previous one is answered from a cached
 * bitmap without being parsed.  Group / pod names are interned to the dense ids of hived.h with recycling.       */
#ifndef HIVED_INGEST_H_
#define HIVED_INGEST_H_
#include "hived.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct hived_ingest hived_ingest;

int hived_ingest_create(hived_ctx*, hived_ingest** out);
void hived_ingest_destroy(hived_ingest*);
int32_t hived_ingest_bitmap_words(const hived_ingest*);            /* (number of nodes + 31) / 32 */
int32_t hived_ingest_node_id(const hived_ingest*, const char* name, int32_t len); /* -1: not a node of the cluster */

/* names[0..n): NUL-terminated node names.  bitmap_out[hived_ingest_bitmap_words()] receives the node set; names that
 * are not nodes of the cluster are skipped (the reference's set would hold them, and no cell would ever look them
 * up).  Returns the number of DISTINCT known nodes; *is_all = 1 when that is every node of the cluster.          */
int32_t hived_ingest_node_names(hived_ingest*, const char* const* names, int32_t n, uint32_t* bitmap_out, int32_t* is_all);

/* The same from JSON text: `json` points at (or before) the '[' of a JSON array of strings — the value of
 * "NodeNames" in schedulerapi.ExtenderArgs — and the array is consumed without materialising any string.
 * *consumed = bytes up to and including the closing ']'.  A body byte-identical to the previous call's is answered
 * from the cache (*cached = 1).  Returns the number of distinct known nodes, or -1 on malformed JSON.              */
int32_t hived_ingest_node_names_json(hived_ingest*, const char* json, int64_t len, uint32_t* bitmap_out, int32_t* is_all,
                                     int64_t* consumed, int32_t* cached);

/* Finds `"key"` at the top level of the JSON object in [json, json+len) and returns the offset of its value (-1: absent). */
int64_t hived_ingest_json_find(const char* json, int64_t len, const char* key);

/* Group / pod names -> dense ids (hived.h "Id lifetime"): intern returns the existing id or the lowest free one
 * (-1: the table of `capacity` ids is full); release frees the id for reuse.  kind: 0 groups, 1 pods.            */
int32_t hived_ingest_intern(hived_ingest*, int32_t kind, const char* name, int32_t len, int32_t capacity);
int32_t hived_ingest_lookup(const hived_ingest*, int32_t kind, const char* name, int32_t len);
int32_t hived_ingest_release(hived_ingest*, int32_t kind, const char* name, int32_t len);

/* The scheduling-spec annotation (api.PodSchedulingSpec as YAML, api/types.go:55-83) -> hived_pod_spec_t.
 * Interns the affinity group's name (group capacity `max_groups`) and `pod_name` (capacity `max_pods`); resolves
 * virtualCluster / leafCellType / pinnedCellId against the scheduler's tables.  Defaults as in
 * ExtractPodSchedulingSpec (internal/utils.go:244-287): no affinityGroup -> a gang of its own named after the pod;
 * leafCellNumber falls back to gpuNumber; the v1 field names (gpuType, gpuNumber, ...: convertOldAnnotation :187-197) are accepted.
 * Returns 0, or HIVED_ERR_BAD_SPEC / HIVED_ERR_UNKNOWN_VC / ... with hived_ingest_last_error() describing why.    */
int hived_ingest_pod_spec_yaml(hived_ingest*, const char* yaml, int64_t len, const char* pod_name, int32_t max_groups,
                               int32_t max_pods, hived_pod_spec_t* out);
const char* hived_ingest_last_error(const hived_ingest*);
const char* hived_ingest_last_group_name(const hived_ingest*);  /* AffinityGroup.Name of the last annotation parsed */
#ifdef __cplusplus
}
#endif
#endif
