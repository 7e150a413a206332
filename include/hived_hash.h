/* FNV-1a parity hash over scheduling results.  ABI-level helper shared by
 * every implementation of hived.h: it only looks at hived_result_t + the caller's pool. */
#ifndef HIVED_HASH_H_
#define HIVED_HASH_H_
#include <stdint.h>

#include "hived.h"

#define HIVED_FNV_OFFSET 0xcbf29ce484222325ull
#define HIVED_FNV_PRIME 0x100000001b3ull

static inline uint64_t hived_fnv_i32(uint64_t h, int32_t v) {
  uint32_t u = (uint32_t)v;
  for (int i = 0; i < 4; i++) {
    h ^= (uint64_t)((u >> (8 * i)) & 0xffu);
    h *= HIVED_FNV_PRIME;
  }
  return h;
}

/* kind; bind: chain, pod_index, node, every leaf triple; preempt: every victim pair;
 * wait: wait_code, wait_cell; errors: the code. */
static inline uint64_t hived_hash_result(uint64_t h, const hived_result_t* r, const int32_t* pool) {
  h = hived_fnv_i32(h, r->error);
  if (r->error != 0) return h;
  h = hived_fnv_i32(h, r->kind);
  if (r->kind == HIVED_KIND_BIND) {
    h = hived_fnv_i32(h, r->chain);
    h = hived_fnv_i32(h, r->pod_index);
    h = hived_fnv_i32(h, r->node);
    h = hived_fnv_i32(h, r->n_members);
    for (int32_t i = 0; i < r->n_members; i++) {
      h = hived_fnv_i32(h, r->member_leaf_num[i]);
      h = hived_fnv_i32(h, r->member_pod_num[i]);
    }
    for (int32_t i = 0; i < 3 * r->n_leaves; i++) h = hived_fnv_i32(h, pool[r->leaf_off + i]);
  } else if (r->kind == HIVED_KIND_PREEMPT) {
    for (int32_t i = 0; i < 2 * r->n_victims; i++) h = hived_fnv_i32(h, pool[r->victim_off + i]);
  } else {
    h = hived_fnv_i32(h, r->wait_code);
    h = hived_fnv_i32(h, r->wait_cell);
  }
  return h;
}
#endif
