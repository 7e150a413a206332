// Request ingest (include/hived_ingest.h): host-side helpers of the C ABI.
// Pure host code — string interning, JSON / YAML scanning — compiled into every backend of hived.h.
//
// Reference behaviour restated (never copied):
//   hived_algorithm.go:190-193   suggestedNodes -> set           -> nodeNames()/nodeNamesJson(): bitmap over node ids
//   webserver.go:173-182         ExtenderArgs JSON decode        -> jsonFind() + nodeNamesJson() on the raw body
//   internal/utils.go:187-197, 230-287  annotation -> PodSchedulingSpec (defaulting + validation) -> podSpecYaml()
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hived.h"
#include "../../include/hived_ingest.h"

namespace hived_ingest_impl {

inline uint64_t fnv1a(const char* s, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ull; }
  return h;
}

// names -> dense ids with recycling (lowest free id first, so that a steady state stays inside a small id range)
struct Interner {
  std::unordered_map<std::string, int32_t> ids;
  std::vector<int32_t> freeIds;  // min-heap
  int32_t next = 0;
  int32_t intern(const char* s, int32_t len, int32_t capacity) {
    std::string k(s, (size_t)len);
    auto it = ids.find(k);
    if (it != ids.end()) return it->second;
    int32_t id;
    if (!freeIds.empty()) {
      std::pop_heap(freeIds.begin(), freeIds.end(), std::greater<int32_t>());
      id = freeIds.back();
      freeIds.pop_back();
    } else {
      if (next >= capacity) return -1;
      id = next++;
    }
    ids.emplace(std::move(k), id);
    return id;
  }
  int32_t lookup(const char* s, int32_t len) const {
    auto it = ids.find(std::string(s, (size_t)len));
    return it == ids.end() ? -1 : it->second;
  }
  int32_t release(const char* s, int32_t len) {
    auto it = ids.find(std::string(s, (size_t)len));
    if (it == ids.end()) return -1;
    int32_t id = it->second;
    ids.erase(it);
    freeIds.push_back(id);
    std::push_heap(freeIds.begin(), freeIds.end(), std::greater<int32_t>());
    return id;
  }
};

// ---- a YAML subset large enough for api.PodSchedulingSpec: block mappings / sequences by indentation, flow
// mappings / sequences ({...}, [...]: JSON included), plain / single- / double-quoted scalars, comments.
struct YNode {
  enum Kind { SCALAR, MAP, SEQ } kind = SCALAR;
  std::string scalar;
  bool quoted = false;
  std::vector<std::pair<std::string, YNode>> map;
  std::vector<YNode> seq;
  const YNode* get(const char* key) const {
    if (kind != MAP) return nullptr;
    for (auto& kv : map) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct YParser {
  const char* p;
  const char* end;
  std::string err;
  struct Line { int indent; const char* b; const char* e; };
  std::vector<Line> lines;
  size_t li = 0;

  static bool isBlank(char c) { return c == ' ' || c == '\t' || c == '\r'; }
  // strip a trailing comment (" #...") outside quotes and trailing blanks
  static const char* trimEnd(const char* b, const char* e) {
    char q = 0;
    for (const char* c = b; c < e; c++) {
      if (q) { if (*c == q) q = 0; else if (q == '"' && *c == '\\' && c + 1 < e) c++; continue; }
      if (*c == '"' || *c == '\'') { q = *c; continue; }
      if (*c == '#' && (c == b || isBlank(c[-1]))) { e = c; break; }
    }
    while (e > b && isBlank(e[-1])) e--;
    return e;
  }
  void splitLines() {
    const char* c = p;
    while (c < end) {
      const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
      const char* le = nl ? nl : end;
      int ind = 0;
      const char* b = c;
      while (b < le && *b == ' ') { b++; ind++; }
      const char* e = trimEnd(b, le);
      if (e > b && !(e - b == 3 && !memcmp(b, "---", 3))) lines.push_back({ind, b, e});
      c = nl ? nl + 1 : end;
    }
  }
  // ---- flow style (and every scalar)
  void skipWs(const char*& c, const char* e) { while (c < e && (isBlank(*c) || *c == '\n')) c++; }
  bool parseQuoted(const char*& c, const char* e, std::string& out) {
    const char q = *c++;
    out.clear();
    while (c < e && *c != q) {
      if (q == '"' && *c == '\\' && c + 1 < e) {
        c++;
        switch (*c) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'u': {  // \uXXXX -> UTF-8
            if (c + 4 >= e) { err = "bad \\u escape"; return false; }
            unsigned v = 0;
            for (int i = 1; i <= 4; i++) {
              char h = c[i];
              v = v * 16 + (h >= '0' && h <= '9' ? h - '0' : (h | 32) >= 'a' && (h | 32) <= 'f' ? (h | 32) - 'a' + 10 : 0);
            }
            c += 4;
            if (v < 0x80) out += (char)v;
            else if (v < 0x800) { out += (char)(0xC0 | (v >> 6)); out += (char)(0x80 | (v & 0x3F)); }
            else { out += (char)(0xE0 | (v >> 12)); out += (char)(0x80 | ((v >> 6) & 0x3F)); out += (char)(0x80 | (v & 0x3F)); }
            break;
          }
          default: out += *c;
        }
        c++;
      } else if (q == '\'' && *c == '\'' && c + 1 < e && c[1] == '\'') { out += '\''; c += 2; }
      else out += *c++;
    }
    if (c >= e) { err = "unterminated quoted scalar"; return false; }
    c++;
    return true;
  }
  bool parseFlow(const char*& c, const char* e, YNode& n, int depth = 0) {
    if (depth > 16) { err = "nesting too deep"; return false; }
    skipWs(c, e);
    if (c >= e) { n.kind = YNode::SCALAR; n.scalar.clear(); return true; }
    if (*c == '{') {
      n.kind = YNode::MAP;
      c++;
      skipWs(c, e);
      if (c < e && *c == '}') { c++; return true; }
      while (true) {
        YNode k;
        skipWs(c, e);
        if (c < e && (*c == '"' || *c == '\'')) { if (!parseQuoted(c, e, k.scalar)) return false; }
        else { const char* b = c; while (c < e && *c != ':' && *c != ',' && *c != '}') c++; const char* ke = c; while (ke > b && isBlank(ke[-1])) ke--; k.scalar.assign(b, ke); }
        skipWs(c, e);
        if (c >= e || *c != ':') { err = "expected ':' in flow mapping"; return false; }
        c++;
        YNode v;
        if (!parseFlow(c, e, v, depth + 1)) return false;
        n.map.emplace_back(std::move(k.scalar), std::move(v));
        skipWs(c, e);
        if (c < e && *c == ',') { c++; continue; }
        if (c < e && *c == '}') { c++; return true; }
        err = "expected ',' or '}' in flow mapping";
        return false;
      }
    }
    if (*c == '[') {
      n.kind = YNode::SEQ;
      c++;
      skipWs(c, e);
      if (c < e && *c == ']') { c++; return true; }
      while (true) {
        YNode v;
        if (!parseFlow(c, e, v, depth + 1)) return false;
        n.seq.push_back(std::move(v));
        skipWs(c, e);
        if (c < e && *c == ',') { c++; continue; }
        if (c < e && *c == ']') { c++; return true; }
        err = "expected ',' or ']' in flow sequence";
        return false;
      }
    }
    n.kind = YNode::SCALAR;
    if (*c == '"' || *c == '\'') { n.quoted = true; return parseQuoted(c, e, n.scalar); }
    const char* b = c;
    while (c < e && *c != ',' && *c != '}' && *c != ']' && *c != '\n') c++;
    const char* se = c;
    while (se > b && isBlank(se[-1])) se--;
    n.scalar.assign(b, se);
    return true;
  }
  // a value that starts on a line: the rest of the line (flow collections may not span lines here)
  bool parseInline(const char* b, const char* e, YNode& n) {
    const char* c = b;
    if (!parseFlow(c, e, n)) return false;
    // a plain scalar may contain ',' / ']' / '}' in block context: take the whole rest of the line then
    if (n.kind == YNode::SCALAR && !n.quoted) { n.scalar.assign(b, e); }
    return true;
  }
  // key of a "key: value" line; returns pointer past the ':' (nullptr: not a mapping line)
  static const char* splitKey(const char* b, const char* e, std::string& key) {
    const char* c = b;
    if (c < e && (*c == '"' || *c == '\'')) {
      const char q = *c++;
      const char* kb = c;
      while (c < e && *c != q) c++;
      if (c >= e) return nullptr;
      key.assign(kb, c);
      c++;
      while (c < e && isBlank(*c)) c++;
      if (c < e && *c == ':') return c + 1;
      return nullptr;
    }
    for (; c < e; c++)
      if (*c == ':' && (c + 1 == e || isBlank(c[1]))) {
        const char* ke = c;
        while (ke > b && isBlank(ke[-1])) ke--;
        key.assign(b, ke);
        return c + 1;
      }
    return nullptr;
  }
  bool parseBlock(int indent, YNode& n, int depth = 0) {
    if (depth > 16) { err = "nesting too deep"; return false; }
    if (li >= lines.size()) { n.kind = YNode::SCALAR; return true; }
    const Line first = lines[li];
    if (first.b[0] == '-' && (first.e - first.b == 1 || isBlank(first.b[1]))) {
      n.kind = YNode::SEQ;
      while (li < lines.size() && lines[li].indent == indent && lines[li].b[0] == '-' &&
             (lines[li].e - lines[li].b == 1 || isBlank(lines[li].b[1]))) {
        Line& L = lines[li];
        const char* c = L.b + 1;
        int extra = 1;
        while (c < L.e && isBlank(*c)) { c++; extra++; }
        YNode item;
        if (c >= L.e) {  // "-" alone: the item is the nested block
          li++;
          if (li < lines.size() && lines[li].indent > indent) { if (!parseBlock(lines[li].indent, item, depth + 1)) return false; }
        } else {
          std::string k;
          if (*c != '{' && *c != '[' && splitKey(c, L.e, k)) {
            // "- key: value": a mapping whose first entry sits on the dash line; rewrite the line as that entry
            L.indent = indent + extra;
            L.b = c;
            if (!parseBlock(L.indent, item, depth + 1)) return false;
          } else {
            if (!parseInline(c, L.e, item)) return false;
            li++;
          }
        }
        n.seq.push_back(std::move(item));
      }
      return true;
    }
    std::string key;
    const char* rest = splitKey(first.b, first.e, key);
    if (!rest) {  // a lone scalar / flow collection
      li++;
      return parseInline(first.b, first.e, n);
    }
    n.kind = YNode::MAP;
    while (li < lines.size() && lines[li].indent == indent) {
      const Line L = lines[li];
      std::string k;
      const char* r = splitKey(L.b, L.e, k);
      if (!r) { err = "expected 'key: value'"; return false; }
      while (r < L.e && isBlank(*r)) r++;
      YNode v;
      li++;
      if (r < L.e) {
        if (!parseInline(r, L.e, v)) return false;
      } else if (li < lines.size() && (lines[li].indent > indent ||
                                       (lines[li].indent == indent && lines[li].b[0] == '-' &&
                                        (lines[li].e - lines[li].b == 1 || isBlank(lines[li].b[1]))))) {
        // nested block; a block sequence may sit at the indentation of its key
        if (!parseBlock(lines[li].indent, v, depth + 1)) return false;
      }
      n.map.emplace_back(std::move(k), std::move(v));
    }
    if (li < lines.size() && lines[li].indent > indent) { err = "bad indentation"; return false; }
    return true;
  }
  bool parse(const char* b, const char* e, YNode& root) {
    p = b; end = e;
    const char* c = b;
    skipWs(c, e);
    if (c < e && (*c == '{' || *c == '[')) {  // whole document in flow style (JSON)
      return parseFlow(c, e, root);
    }
    splitLines();
    if (lines.empty()) { root.kind = YNode::SCALAR; return true; }
    return parseBlock(lines[0].indent, root);
  }
};

inline bool yInt(const YNode* n, long long& out) {
  if (!n || n->kind != YNode::SCALAR || n->scalar.empty()) return false;
  char* endp = nullptr;
  out = strtoll(n->scalar.c_str(), &endp, 10);
  return endp && *endp == 0;
}
inline bool yBool(const YNode* n, bool& out) {
  if (!n || n->kind != YNode::SCALAR) return false;
  std::string s = n->scalar;
  for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
  if (s == "true" || s == "yes" || s == "on" || s == "y") { out = true; return true; }
  if (s == "false" || s == "no" || s == "off" || s == "n") { out = false; return true; }
  return false;
}

}  // namespace hived_ingest_impl

struct hived_ingest {
  hived_ctx* ctx = nullptr;
  int32_t nNodes = 0, words = 0;
  // node names: open addressing, (hash, id) slots
  std::vector<std::string> names;
  std::vector<uint64_t> slotHash;
  std::vector<int32_t> slotId;
  uint64_t mask = 0;
  hived_ingest_impl::Interner groups, pods;
  std::unordered_map<std::string, int32_t> vcIds, pinnedIds, leafTypeIds;
  // the previous request's NodeNames array, verbatim, and what it decoded to
  std::string cachedText;
  std::vector<uint32_t> cachedBitmap;
  int32_t cachedCount = -1;
  std::string err, lastGroup;

  void build() {
    size_t cap = 16;
    while (cap < (size_t)nNodes * 2) cap <<= 1;
    slotHash.assign(cap, 0);
    slotId.assign(cap, -1);
    mask = cap - 1;
    for (int32_t i = 0; i < nNodes; i++) {
      uint64_t h = hived_ingest_impl::fnv1a(names[i].data(), names[i].size());
      size_t s = (size_t)(h & mask);
      while (slotId[s] >= 0) s = (s + 1) & mask;
      slotHash[s] = h;
      slotId[s] = i;
    }
  }
  int32_t findHashed(uint64_t h, const char* s, size_t n) const {
    size_t k = (size_t)(h & mask);
    while (slotId[k] >= 0) {
      if (slotHash[k] == h) {
        const std::string& nm = names[slotId[k]];
        if (nm.size() == n && !memcmp(nm.data(), s, n)) return slotId[k];
      }
      k = (k + 1) & mask;
    }
    return -1;
  }
  int32_t find(const char* s, size_t n) const { return findHashed(hived_ingest_impl::fnv1a(s, n), s, n); }
};

extern "C" {

int hived_ingest_create(hived_ctx* ctx, hived_ingest** out) {
  *out = nullptr;
  if (!ctx) return HIVED_ERR_BAD_SPEC;
  hived_ingest* g = new hived_ingest();
  g->ctx = ctx;
  g->nNodes = hived_num_nodes(ctx);
  g->words = (g->nNodes + 31) / 32;
  g->names.reserve((size_t)g->nNodes);
  for (int32_t i = 0; i < g->nNodes; i++) { const char* s = hived_node_name(ctx, i); g->names.emplace_back(s ? s : ""); }
  g->build();
  for (int32_t i = 0; i < hived_num_vcs(ctx); i++) g->vcIds[hived_vc_name(ctx, i)] = i;
  for (int32_t i = 0; i < hived_num_pinned(ctx); i++) g->pinnedIds[hived_pinned_name(ctx, i)] = i;
  for (int32_t i = 0; i < hived_num_leaf_types(ctx); i++) g->leafTypeIds[hived_leaf_type_name(ctx, i)] = i;
  *out = g;
  return 0;
}
void hived_ingest_destroy(hived_ingest* g) { delete g; }
int32_t hived_ingest_bitmap_words(const hived_ingest* g) { return g->words; }
int32_t hived_ingest_node_id(const hived_ingest* g, const char* name, int32_t len) {
  return g->find(name, len < 0 ? strlen(name) : (size_t)len);
}
const char* hived_ingest_last_error(const hived_ingest* g) { return g->err.c_str(); }
const char* hived_ingest_last_group_name(const hived_ingest* g) { return g->lastGroup.c_str(); }

int32_t hived_ingest_node_names(hived_ingest* g, const char* const* names, int32_t n, uint32_t* bm, int32_t* is_all) {
  memset(bm, 0, (size_t)g->words * 4);
  int32_t cnt = 0;
  for (int32_t i = 0; i < n; i++) {
    const char* s = names[i];
    if (!s) continue;
    int32_t id = g->find(s, strlen(s));
    if (id < 0) continue;
    uint32_t& w = bm[id >> 5];
    const uint32_t bit = 1u << (id & 31);
    if (!(w & bit)) { w |= bit; cnt++; }
  }
  if (is_all) *is_all = cnt == g->nNodes ? 1 : 0;
  return cnt;
}

int32_t hived_ingest_node_names_json(hived_ingest* g, const char* json, int64_t len, uint32_t* bm, int32_t* is_all,
                                            int64_t* consumed, int32_t* cached) {
  const char* c = json;
  const char* e = json + len;
  if (cached) *cached = 0;
  while (c < e && *c != '[') {
    if (*c != ' ' && *c != '\t' && *c != '\n' && *c != '\r' && *c != ':') { g->err = "NodeNames: expected a JSON array"; return -1; }
    c++;
  }
  if (c >= e) { g->err = "NodeNames: expected a JSON array"; return -1; }
  const char* arr = c;
  // the previous request's array, byte for byte (it ends with its ']': a match is a whole array)
  if (g->cachedCount >= 0 && (size_t)(e - arr) >= g->cachedText.size() && !memcmp(arr, g->cachedText.data(), g->cachedText.size())) {
    memcpy(bm, g->cachedBitmap.data(), (size_t)g->words * 4);
    if (is_all) *is_all = g->cachedCount == g->nNodes ? 1 : 0;
    if (consumed) *consumed = (int64_t)(arr - json) + (int64_t)g->cachedText.size();
    if (cached) *cached = 1;
    return g->cachedCount;
  }
  memset(bm, 0, (size_t)g->words * 4);
  int32_t cnt = 0;
  c++;
  std::string tmp;
  while (true) {
    while (c < e && (*c == ' ' || *c == '\t' || *c == '\n' || *c == '\r' || *c == ',')) c++;
    if (c >= e) { g->err = "NodeNames: unterminated array"; return -1; }
    if (*c == ']') { c++; break; }
    if (*c != '"') { g->err = "NodeNames: expected a string"; return -1; }
    c++;
    const char* s = c;
    uint64_t h = 0xcbf29ce484222325ull;
    bool esc = false;
    while (c < e && *c != '"') {
      if (*c == '\\') { esc = true; break; }
      h ^= (unsigned char)*c; h *= 0x100000001b3ull;
      c++;
    }
    int32_t id;
    if (esc) {  // rare: decode the escapes into a scratch string
      hived_ingest_impl::YParser yp;
      const char* q = s - 1;
      if (!yp.parseQuoted(q, e, tmp)) { g->err = "NodeNames: " + yp.err; return -1; }
      c = q;
      id = g->find(tmp.data(), tmp.size());
    } else {
      if (c >= e) { g->err = "NodeNames: unterminated string"; return -1; }
      id = g->findHashed(h, s, (size_t)(c - s));
      c++;
    }
    if (id >= 0) {
      uint32_t& w = bm[id >> 5];
      const uint32_t bit = 1u << (id & 31);
      if (!(w & bit)) { w |= bit; cnt++; }
    }
  }
  g->cachedText.assign(arr, (size_t)(c - arr));
  g->cachedBitmap.assign(bm, bm + g->words);
  g->cachedCount = cnt;
  if (is_all) *is_all = cnt == g->nNodes ? 1 : 0;
  if (consumed) *consumed = (int64_t)(c - json);
  return cnt;
}

int64_t hived_ingest_json_find(const char* json, int64_t len, const char* key) {
  const char* c = json;
  const char* e = json + len;
  const size_t kl = strlen(key);
  int depth = 0;
  while (c < e) {
    const char ch = *c;
    if (ch == '"') {
      const char* s = ++c;
      while (c < e && *c != '"') { if (*c == '\\' && c + 1 < e) c++; c++; }
      const char* se = c;
      if (c < e) c++;
      if (depth == 1 && (size_t)(se - s) == kl && !memcmp(s, key, kl)) {
        const char* v = c;
        while (v < e && (*v == ' ' || *v == '\t' || *v == '\n' || *v == '\r')) v++;
        if (v < e && *v == ':') {
          v++;
          while (v < e && (*v == ' ' || *v == '\t' || *v == '\n' || *v == '\r')) v++;
          return (int64_t)(v - json);
        }
      }
      continue;
    }
    if (ch == '{' || ch == '[') depth++;
    else if (ch == '}' || ch == ']') depth--;
    c++;
  }
  return -1;
}

int32_t hived_ingest_intern(hived_ingest* g, int32_t kind, const char* name, int32_t len, int32_t capacity) {
  return (kind ? g->pods : g->groups).intern(name, len < 0 ? (int32_t)strlen(name) : len, capacity);
}
int32_t hived_ingest_lookup(const hived_ingest* g, int32_t kind, const char* name, int32_t len) {
  return (kind ? g->pods : g->groups).lookup(name, len < 0 ? (int32_t)strlen(name) : len);
}
int32_t hived_ingest_release(hived_ingest* g, int32_t kind, const char* name, int32_t len) {
  return (kind ? g->pods : g->groups).release(name, len < 0 ? (int32_t)strlen(name) : len);
}

int hived_ingest_pod_spec_yaml(hived_ingest* g, const char* yaml, int64_t len, const char* pod_name, int32_t max_groups,
                                      int32_t max_pods, hived_pod_spec_t* out) {
  using namespace hived_ingest_impl;
  const char* pfx = "Pod annotation hivedscheduler.microsoft.com/pod-scheduling-spec: ";
  auto bad = [&](const std::string& m) { g->err = pfx + m; return (int)HIVED_ERR_BAD_SPEC; };
  memset(out, 0, sizeof *out);
  if (!yaml || len <= 0) return bad("Annotation does not exist or is empty");
  // convertOldAnnotation (internal/utils.go:187-197): the v1 field names, replaced textually like the reference does
  std::string text(yaml, (size_t)len);
  auto replaceAll = [&](const char* from, const char* to) {
    const size_t fl = strlen(from), tl = strlen(to);
    for (size_t pos = 0; (pos = text.find(from, pos)) != std::string::npos; pos += tl) text.replace(pos, fl, to);
  };
  replaceAll("gpuType", "leafCellType");
  replaceAll("gpuNumber", "leafCellNumber");
  replaceAll("gpuIsolation", "leafCellIsolation");
  replaceAll("physicalGpuIndices", "physicalLeafCellIndices");
  bool onlyWs = true;
  for (char ch : text) if (ch != ' ' && ch != '\n' && ch != '\t' && ch != '\r') { onlyWs = false; break; }
  if (onlyWs) return bad("Annotation does not exist or is empty");
  YParser yp;
  YNode root;
  if (!yp.parse(text.data(), text.data() + text.size(), root)) return bad("cannot parse: " + yp.err);
  if (root.kind != YNode::MAP) return bad("cannot parse: not a mapping");
  auto str = [&](const char* k) { const YNode* n = root.get(k); return (n && n->kind == YNode::SCALAR) ? n->scalar : std::string(); };
  const std::string vc = str("virtualCluster"), pinned = str("pinnedCellId"), leafType = str("leafCellType");
  long long priority = 0, leafNum = 0;
  if (root.get("priority") && !yInt(root.get("priority"), priority)) return bad("cannot parse: priority is not an integer");
  if (root.get("leafCellNumber") && !yInt(root.get("leafCellNumber"), leafNum)) return bad("cannot parse: leafCellNumber is not an integer");
  bool lazy = false, ignoreSuggested = true;  // IgnoreK8sSuggestedNodes defaults to true (internal/utils.go:236)
  if (root.get("lazyPreemptionEnable") && !yBool(root.get("lazyPreemptionEnable"), lazy)) return bad("cannot parse: lazyPreemptionEnable is not a bool");
  if (root.get("ignoreK8sSuggestedNodes") && !yBool(root.get("ignoreK8sSuggestedNodes"), ignoreSuggested))
    return bad("cannot parse: ignoreK8sSuggestedNodes is not a bool");
  std::string groupName;
  std::vector<std::pair<long long, long long>> members;  // (podNumber, leafCellNumber)
  const YNode* ag = root.get("affinityGroup");
  if (ag && ag->kind == YNode::MAP) {
    const YNode* nm = ag->get("name");
    if (nm && nm->kind == YNode::SCALAR) groupName = nm->scalar;
    const YNode* ms = ag->get("members");
    if (ms && ms->kind == YNode::SEQ)
      for (auto& m : ms->seq) {
        long long pn = 0, ln = 0;
        if (m.kind != YNode::MAP) return bad("cannot parse: affinityGroup.members entry is not a mapping");
        if (m.get("podNumber") && !yInt(m.get("podNumber"), pn)) return bad("cannot parse: podNumber is not an integer");
        if (m.get("leafCellNumber") && !yInt(m.get("leafCellNumber"), ln)) return bad("cannot parse: leafCellNumber is not an integer");
        members.emplace_back(pn, ln);
      }
  } else {  // defaulting: a gang of its own, named after the pod
    groupName = pod_name ? pod_name : "";
    members.emplace_back(1, leafNum);
  }
  // validation, in the reference's order
  if (vc.empty()) return bad("VirtualCluster is empty");
  if (priority < -1) return bad("Priority is less than -1");
  if (priority > HIVED_MAX_GUARANTEED_PRIORITY) return bad("Priority is greater than 1000");
  if (leafNum <= 0) return bad("LeafCellNumber is non-positive");
  if (groupName.empty()) return bad("AffinityGroup.Name is empty");
  bool inGroup = false;
  for (auto& m : members) {
    if (m.first <= 0) return bad("AffinityGroup.Members has non-positive PodNumber");
    if (m.second <= 0) return bad("AffinityGroup.Members has non-positive LeafCellNumber");
    if (m.second == leafNum) inGroup = true;
  }
  if (!inGroup) return bad("AffinityGroup.Members does not contains current Pod");
  if (members.size() > HIVED_MAX_MEMBERS) { g->err = "affinity group has more members than HIVED_MAX_MEMBERS"; return HIVED_ERR_CAPACITY; }
  out->pod = pod_name ? g->pods.intern(pod_name, (int32_t)strlen(pod_name), max_pods) : -1;
  out->group = g->groups.intern(groupName.data(), (int32_t)groupName.size(), max_groups);
  g->lastGroup = groupName;
  if (out->pod < 0 || out->group < 0) { g->err = "id table full (max_groups / max_pods)"; return HIVED_ERR_CAPACITY; }
  auto idOf = [](const std::unordered_map<std::string, int32_t>& m, const std::string& k, int32_t none, int32_t unknown) {
    if (k.empty()) return none;
    auto it = m.find(k);
    return it == m.end() ? unknown : it->second;
  };
  out->vc = idOf(g->vcIds, vc, -1, -1);
  out->priority = (int32_t)priority;
  out->pinned = idOf(g->pinnedIds, pinned, -1, -2);
  out->leaf_type = idOf(g->leafTypeIds, leafType, -1, -2);
  out->leaf_num = (int32_t)leafNum;
  out->flags = (lazy ? HIVED_SPEC_LAZY_PREEMPTION : 0) | (ignoreSuggested ? HIVED_SPEC_IGNORE_SUGGESTED : 0);
  out->n_members = (int32_t)members.size();
  for (size_t i = 0; i < members.size(); i++) { out->member_pod_num[i] = (int32_t)members[i].first; out->member_leaf_num[i] = (int32_t)members[i].second; }
  g->err.clear();
  return 0;
}

}  // extern "C"
