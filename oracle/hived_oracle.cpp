// TEST INFRASTRUCTURE — CPU oracle (see hived_oracle.hpp).  Restates microsoft/hivedscheduler
// pkg/algorithm/*.go function by function; each function cites the reference lines it follows.
#include "hived_oracle.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <sstream>

#include "../include/hived.h"

namespace hived_oracle {

// ------------------------------------------------------------------------------------------
// small helpers mirroring types.go / common
// ------------------------------------------------------------------------------------------

// cell.go:50-56
static inline bool CellEqual(const Cell* c1, const Cell* c2) {
  if (c1 == nullptr || c2 == nullptr) return c1 == nullptr && c2 == nullptr;
  return c1->address == c2->address;
}

// types.go:69-76
static bool listContains(const CellList& cl, const Cell* c) {
  for (const Cell* cc : cl)
    if (CellEqual(cc, c)) return true;
  return false;
}

// types.go:78-95 — swap-with-last removal
static void listRemove(CellList& cl, const Cell* c) {
  int index = -1;
  for (size_t i = 0; i < cl.size(); i++) {
    if (CellEqual(cl[i], c)) {
      index = (int)i;
      break;
    }
  }
  if (index < 0) throw Panic("Cell not not found in list when removing: " + c->address);
  cl[index] = cl.back();
  cl.pop_back();
}

// types.go:123-130
static ChainCellList shallowCopy(const ChainCellList& ccl) {
  ChainCellList copied;
  for (int32_t l = 1; l <= ccl.len(); l++) copied.m[l] = ccl.at(l);
  return copied;
}

static std::vector<std::string> splitStr(const std::string& s, char sep) {
  std::vector<std::string> out;
  size_t start = 0;
  while (true) {
    size_t p = s.find(sep, start);
    if (p == std::string::npos) {
      out.push_back(s.substr(start));
      break;
    }
    out.push_back(s.substr(start, p - start));
    start = p + 1;
  }
  return out;
}

// common.StringToInt32 (pkg/common/utils.go:248-254): panics on malformed input
static int32_t StringToInt32(const std::string& s) {
  char* end = nullptr;
  long v = strtol(s.c_str(), &end, 10);
  if (s.empty() || *end != '\0') throw Panic("strconv.ParseInt: parsing \"" + s + "\": invalid syntax", HIVED_ERR_BAD_CONFIG);
  return (int32_t)v;
}

static int32_t usedAt(const Cell* c, int32_t p) {
  auto it = c->used.find(p);
  return it == c->used.end() ? 0 : it->second;
}

// cell.go:122-127
static void IncreaseUsedLeafCellNumAtPriority(Cell* c, int32_t p, int32_t delta) {
  c->used[p] += delta;
  if (c->used[p] == 0) c->used.erase(p);
}

// cell.go:195-204
static void SetState(Cell* c, int32_t s) {
  c->state = s;
  if (c->virtualCell != nullptr) c->virtualCell->state = s;
}

// cell.go:302-312
static void SetHealthiness(Cell* c, bool healthy) {
  c->healthy = healthy;
  if (c->virtualCell != nullptr) c->virtualCell->healthy = c->healthy;
}

// cell.go:264-277
static void SetVirtualCell(Cell* pc, Cell* vc) { pc->virtualCell = vc; }

// cell.go:401-419
static void SetPhysicalCell(Cell* vc, Cell* pc) {
  vc->physicalCell = pc;
  if (pc == nullptr) {
    vc->state = cellFree;
    vc->healthy = true;
  } else {
    vc->healthy = pc->healthy;
  }
}

// ------------------------------------------------------------------------------------------
// config.go — ParseConfig and the three constructors
// ------------------------------------------------------------------------------------------

namespace {

// config.go:34-43
struct cellChainElement {
  std::string cellType;
  int32_t level = 0;
  std::string childCellType;
  int32_t childNumber = 0;
  bool hasNode = false;
  bool isMultiNodes = false;
  std::string leafCellType;
  int32_t leafCellNumber = 0;
};

struct CellTypeSpec {
  std::string childCellType;
  int32_t childCellNumber = 0;
  bool isNodeLevel = false;
};

struct PhysicalCellSpec {
  std::string cellType, cellAddress, pinnedCellId;
  std::vector<PhysicalCellSpec> children;
};

struct VirtualClusterSpec {
  std::vector<std::pair<std::string, int32_t>> virtualCells;  // (cellType path, number)
  std::vector<std::string> pinnedCells;
};

struct Config {
  std::map<std::string, CellTypeSpec> cellTypes;
  std::vector<PhysicalCellSpec> physicalCells;
  std::map<std::string, VirtualClusterSpec> virtualClusters;
};

struct Tokenizer {
  std::istringstream in;
  explicit Tokenizer(const std::string& s) : in(s) {}
  std::string next() {
    std::string t;
    if (!(in >> t)) throw Panic("HIVEDSPEC: unexpected end of spec", HIVED_ERR_BAD_CONFIG);
    return t;
  }
  int32_t nextInt() { return StringToInt32(next()); }
  void expect(const char* w) {
    std::string t = next();
    if (t != w) throw Panic(std::string("HIVEDSPEC: expected '") + w + "' got '" + t + "'", HIVED_ERR_BAD_CONFIG);
  }
};

static void readPhysical(Tokenizer& tk, PhysicalCellSpec& spec, int depth) {
  int32_t d = tk.nextInt();
  if (d != depth) throw Panic("HIVEDSPEC: bad physical cell depth", HIVED_ERR_BAD_CONFIG);
  spec.cellType = tk.next();
  spec.cellAddress = tk.next();
  spec.pinnedCellId = tk.next();
  if (spec.pinnedCellId == "-") spec.pinnedCellId.clear();
  int32_t n = tk.nextInt();
  spec.children.resize(n);
  for (int32_t i = 0; i < n; i++) readPhysical(tk, spec.children[i], depth + 1);
}

static Config readSpec(const std::string& text) {
  Config c;
  Tokenizer tk(text);
  tk.expect("HIVEDSPEC");
  tk.expect("1");
  tk.expect("celltypes");
  int32_t n = tk.nextInt();
  for (int32_t i = 0; i < n; i++) {
    std::string name = tk.next();
    CellTypeSpec s;
    s.childCellType = tk.next();
    s.childCellNumber = tk.nextInt();
    s.isNodeLevel = tk.nextInt() != 0;
    c.cellTypes[name] = s;
  }
  tk.expect("physicalcells");
  n = tk.nextInt();
  c.physicalCells.resize(n);
  for (int32_t i = 0; i < n; i++) readPhysical(tk, c.physicalCells[i], 0);
  tk.expect("virtualclusters");
  n = tk.nextInt();
  for (int32_t i = 0; i < n; i++) {
    tk.expect("vc");
    std::string name = tk.next();
    int32_t nv = tk.nextInt(), np = tk.nextInt();
    VirtualClusterSpec v;
    for (int32_t j = 0; j < nv; j++) {
      std::string t = tk.next();
      int32_t num = tk.nextInt();
      v.virtualCells.push_back({t, num});
    }
    for (int32_t j = 0; j < np; j++) v.pinnedCells.push_back(tk.next());
    c.virtualClusters[name] = v;
  }
  tk.expect("end");
  return c;
}

// config.go:45-109
struct cellTypeConstructor {
  const std::map<std::string, CellTypeSpec>& cellTypeSpecs;
  std::map<std::string, cellChainElement> cellChainElements;
  explicit cellTypeConstructor(const std::map<std::string, CellTypeSpec>& s) : cellTypeSpecs(s) {}

  void addCellChain(const std::string& ct) {  // config.go:59-101
    if (cellChainElements.count(ct)) return;
    auto it = cellTypeSpecs.find(ct);
    if (it == cellTypeSpecs.end()) {
      cellChainElement e;
      e.cellType = ct;
      e.level = lowestLevel;
      e.leafCellType = ct;
      e.leafCellNumber = 1;
      cellChainElements[ct] = e;
      return;
    }
    const CellTypeSpec& ctSpec = it->second;
    const std::string& child = ctSpec.childCellType;
    if (!cellChainElements.count(child)) addCellChain(child);
    const cellChainElement cct = cellChainElements[child];
    cellChainElement e;
    e.cellType = ct;
    e.level = cct.level + 1;
    e.childCellType = cct.cellType;
    e.childNumber = ctSpec.childCellNumber;
    e.hasNode = cct.hasNode || ctSpec.isNodeLevel;
    e.isMultiNodes = cct.hasNode;
    e.leafCellType = cct.leafCellType;
    e.leafCellNumber = cct.leafCellNumber * ctSpec.childCellNumber;
    cellChainElements[ct] = e;
  }
  void buildCellChains() {  // config.go:103-109
    for (auto& kv : cellTypeSpecs) addCellChain(kv.first);
  }
};

}  // namespace

struct HivedAlgorithm::Parsed {
  std::map<std::string, cellChainElement> elements;
  std::map<std::string, Cell*> rawPinnedPhysical;                               // pid -> cell
  std::map<std::string, std::map<std::string, ChainCellList>> nonPinnedFullList;  // vc -> chain
  std::map<std::string, std::map<std::string, ChainCellList>> nonPinnedFreeList;
  std::map<std::string, std::map<std::string, ChainCellList>> pinnedList;         // vc -> pid
  std::map<std::string, std::map<std::string, Cell*>> pinnedPhysicalList;         // vc -> pid
};

void HivedAlgorithm::parseConfig(const std::string& specText) {
  Config cfg = readSpec(specText);
  cellTypeConstructor ctc(cfg.cellTypes);
  ctc.buildCellChains();
  Parsed P;
  P.elements = ctc.cellChainElements;
  auto& elements = P.elements;

  // ---- physicalCellConstructor (config.go:111-246)
  std::string buildingChain;
  auto addCellP = [&](const cellChainElement& ce, const std::string& pid, const std::string& address) -> Cell* {
    // config.go:185-203 + NewPhysicalCell cell.go:144-176
    auto up = std::make_unique<Cell>();
    Cell* c = up.get();
    cellStore_.push_back(std::move(up));
    c->physical = true;
    c->chain = buildingChain;
    c->level = ce.level;
    c->atOrHigherThanNode = ce.hasNode;
    c->totalLeafCellNum = ce.leafCellNumber;
    c->cellType = ce.cellType;
    c->address = address;
    c->isNodeLevel = ce.hasNode && !ce.isMultiNodes;
    fullCellList[buildingChain].mut(ce.level).push_back(c);
    if (!pid.empty()) {
      P.rawPinnedPhysical[pid] = c;
      c->pinned = true;
    }
    return c;
  };
  std::function<Cell*(const PhysicalCellSpec&, const std::string&, std::string)> buildChildP =
      [&](const PhysicalCellSpec& spec, const std::string& ct, std::string currentNode) -> Cell* {
    // config.go:141-183
    auto eit = elements.find(ct);
    if (eit == elements.end()) throw Panic("cellType " + ct + " not found in cell types definition", HIVED_ERR_BAD_CONFIG);
    const cellChainElement& ce = eit->second;
    std::vector<std::string> splitAddress = splitStr(spec.cellAddress, '/');
    if (ce.hasNode && !ce.isMultiNodes) currentNode = splitAddress.back();
    Cell* cellInstance = addCellP(ce, spec.pinnedCellId, spec.cellAddress);
    if (ce.level == 1) {
      cellInstance->nodes = {currentNode};
      cellInstance->leafCellIndices = {StringToInt32(splitAddress.back())};
      return cellInstance;
    }
    std::vector<std::string> currentCellNodes;
    std::vector<int32_t> currentCellLeafCellIndices;
    CellList currentCellChildren;
    for (const PhysicalCellSpec& childSpec : spec.children) {
      Cell* child = buildChildP(childSpec, ce.childCellType, currentNode);
      child->parent = cellInstance;
      currentCellChildren.push_back(child);
      if (ce.isMultiNodes) {
        currentCellNodes.insert(currentCellNodes.end(), child->nodes.begin(), child->nodes.end());
      } else {
        currentCellLeafCellIndices.insert(currentCellLeafCellIndices.end(),
                                          child->leafCellIndices.begin(), child->leafCellIndices.end());
      }
    }
    cellInstance->children = currentCellChildren;
    if (ce.isMultiNodes) {
      currentCellLeafCellIndices = {-1};
    } else {
      currentCellNodes = {currentNode};
    }
    cellInstance->nodes = currentCellNodes;
    cellInstance->leafCellIndices = currentCellLeafCellIndices;
    return cellInstance;
  };
  for (const PhysicalCellSpec& spec : cfg.physicalCells) {  // config.go:231-246
    buildingChain = spec.cellType;
    auto eit = elements.find(buildingChain);  // buildFullTree config.go:216-229
    if (eit == elements.end())
      throw Panic("cellType " + buildingChain + " in PhysicalCells is not found in cell types definition", HIVED_ERR_BAD_CONFIG);
    if (!eit->second.hasNode) throw Panic("top cell must be node-level or above: " + buildingChain, HIVED_ERR_BAD_CONFIG);
    Cell* rootCell = buildChildP(spec, buildingChain, "");
    if (!freeCellList.count(rootCell->chain)) {
      ChainCellList ccl;  // NewChainCellList types.go:101-107
      for (int32_t i = 1; i <= rootCell->level; i++) ccl.m[i] = CellList{};
      freeCellList[rootCell->chain] = ccl;
    }
    freeCellList[rootCell->chain].mut(rootCell->level).push_back(rootCell);
  }

  // ---- virtualCellConstructor (config.go:248-413)
  std::string buildingVc, buildingVChain, buildingChild, buildingPId;
  Cell* buildingRoot = nullptr;
  auto addCellV = [&](const cellChainElement& ce, const std::string& address) -> Cell* {
    // config.go:295-330 + NewVirtualCell cell.go:326-365
    auto up = std::make_unique<Cell>();
    Cell* c = up.get();
    cellStore_.push_back(std::move(up));
    c->physical = false;
    c->vc = buildingVc;
    c->chain = buildingVChain;
    c->level = ce.level;
    c->atOrHigherThanNode = ce.hasNode;
    c->totalLeafCellNum = ce.leafCellNumber;
    c->cellType = ce.cellType;
    c->address = address;
    c->isNodeLevel = ce.hasNode && !ce.isMultiNodes;
    if (buildingPId.empty()) {
      P.nonPinnedFullList[buildingVc][buildingVChain].mut(ce.level).push_back(c);
    } else {
      P.pinnedList[buildingVc][buildingPId].mut(ce.level).push_back(c);
      c->pid = buildingPId;
    }
    if (buildingRoot == nullptr) buildingRoot = c;
    c->preassignedCell = buildingRoot;
    return c;
  };
  std::function<Cell*(const std::string&, const std::string&)> buildChildV =
      [&](const std::string& ct, const std::string& address) -> Cell* {
    // config.go:332-352
    const cellChainElement& ce = elements.at(ct);
    Cell* cellInstance = addCellV(ce, address);
    if (ce.level == 1) return cellInstance;
    CellList currentCellChildren;
    std::vector<std::string> splitAddress = splitStr(address, '/');
    int32_t offset;
    if (splitAddress.size() == 2) {
      offset = 0;
    } else {
      offset = StringToInt32(splitAddress.back()) * ce.childNumber;
    }
    for (int32_t i = 0; i < ce.childNumber; i++) {
      Cell* child = buildChildV(ce.childCellType, address + "/" + std::to_string(offset + i));
      child->parent = cellInstance;
      currentCellChildren.push_back(child);
    }
    cellInstance->children = currentCellChildren;
    return cellInstance;
  };
  for (auto& kv : cfg.virtualClusters) {  // config.go:366-413 (VC order is irrelevant: VCs are independent here)
    const std::string& vc = kv.first;
    const VirtualClusterSpec& spec = kv.second;
    vcFreeCellNum[vc];
    P.nonPinnedFullList[vc];
    P.nonPinnedFreeList[vc];
    P.pinnedList[vc];
    P.pinnedPhysicalList[vc];
    int32_t numCells = 0;
    for (auto& virtualCell : spec.virtualCells) {
      std::vector<std::string> sl = splitStr(virtualCell.first, '.');
      std::string chain = sl[0];
      std::string rootType = sl.back();
      auto rit = elements.find(rootType);
      if (rit == elements.end())
        throw Panic("cellType " + rootType + " in VirtualCells is not found in cell types definition", HIVED_ERR_BAD_CONFIG);
      int32_t rootLevel = rit->second.level;
      vcFreeCellNum[vc][chain][rootLevel] += virtualCell.second;
      for (int32_t i = 0; i < virtualCell.second; i++) {
        buildingVc = vc;
        buildingVChain = chain;
        buildingChild = rootType;
        buildingRoot = nullptr;
        buildingPId.clear();
        Cell* rootCell = buildChildV(buildingChild, vc + "/" + std::to_string(numCells));
        P.nonPinnedFreeList[vc][rootCell->chain].mut(rootCell->level).push_back(rootCell);
        numCells++;
      }
    }
    for (const std::string& pid : spec.pinnedCells) {
      auto pit = P.rawPinnedPhysical.find(pid);
      if (pit == P.rawPinnedPhysical.end())
        throw Panic("pinned cell not found in physicalCells: VC: " + vc + ", ID: " + pid, HIVED_ERR_BAD_CONFIG);
      Cell* pc = pit->second;
      P.pinnedPhysicalList[vc][pid] = pc;
      std::string child = pc->chain;
      while (elements.at(child).level > pc->level) child = elements.at(child).childCellType;
      vcFreeCellNum[vc][pc->chain][pc->level]++;
      buildingVc = vc;
      buildingVChain = pc->chain;
      buildingChild = child;
      buildingRoot = nullptr;
      buildingPId = pid;
      buildChildV(buildingChild, vc + "/" + std::to_string(numCells));
      numCells++;
    }
  }

  // ---- parseCellChainInfo (config.go:415-440).  The chains slice comes from a Go map
  // (config.go:470-473); canonical order = descending name, the order the reference's test pins
  // with sortChains (hived_algorithm_test.go:634-643).
  std::vector<std::string> chains;
  for (auto& kv : fullCellList) chains.push_back(kv.first);
  std::sort(chains.begin(), chains.end(), [](const std::string& a, const std::string& b) { return a > b; });
  for (const std::string& chain : chains) {
    const cellChainElement* ce = &elements.at(chain);
    cellChains[ce->leafCellType].push_back(chain);
    while (true) {
      leafCellNums[chain][ce->level] = ce->leafCellNumber;
      cellTypes[chain][ce->level] = ce->cellType;
      auto nit = elements.find(ce->childCellType);
      if (nit == elements.end()) break;
      ce = &nit->second;
    }
  }

  // ---- NewHivedAlgorithm (hived_algorithm.go:108-145)
  for (auto& kv : P.nonPinnedFullList) {
    const std::string& vcName = kv.first;
    auto up = std::make_unique<IntraVCScheduler>();  // newDefaultIntraVCScheduler intra_vc_scheduler.go:57-78
    IntraVCScheduler* s = up.get();
    vcsStore_.push_back(std::move(up));
    s->nonPinnedFullCellList = P.nonPinnedFullList[vcName];
    s->nonPinnedPreassignedCells = P.nonPinnedFreeList[vcName];
    s->pinnedCells = P.pinnedList[vcName];
    for (auto& cc : s->nonPinnedFullCellList)
      s->nonPinnedCellSchedulers[cc.first] = newTopologyAwareScheduler(cc.second, leafCellNums[cc.first], true);
    for (auto& pc : s->pinnedCells)
      s->pinnedCellSchedulers[pc.first] =
          newTopologyAwareScheduler(pc.second, leafCellNums[pc.second.at(1)[0]->chain], true);
    vcSchedulers[vcName] = s;
  }
  for (auto& kv : fullCellList)
    opportunisticSchedulers[kv.first] = newTopologyAwareScheduler(kv.second, leafCellNums[kv.first], false);

  // interning tables + ABI ids (not in the reference)
  {
    std::set<std::string> types, leafTypes, pinned;
    for (auto& kv : cfg.cellTypes) types.insert(kv.first);
    for (auto& kv : elements) {
      types.insert(kv.first);
      if (kv.second.level == 1) leafTypes.insert(kv.first);
    }
    // only leaf types that occur in a physical chain are schedulable, but ids cover all of them
    cellTypeNames.assign(types.begin(), types.end());
    leafTypeNames.assign(leafTypes.begin(), leafTypes.end());
    for (auto& kv : P.rawPinnedPhysical) pinned.insert(kv.first);
    pinnedNames.assign(pinned.begin(), pinned.end());
    for (auto& kv : fullCellList) chainNames.push_back(kv.first);
    for (auto& kv : cfg.virtualClusters) vcNames.push_back(kv.first);
    std::function<void(const PhysicalCellSpec&, const std::string&)> walk =
        [&](const PhysicalCellSpec& spec, const std::string& ct) {
      auto eit = elements.find(ct);
      const cellChainElement& ce = eit->second;
      if (ce.hasNode && !ce.isMultiNodes) {
        std::string n = splitStr(spec.cellAddress, '/').back();
        if (!nodeIds.count(n)) {
          nodeIds[n] = (int32_t)nodeNames.size();
          nodeNames.push_back(n);
        }
        return;
      }
      for (auto& ch : spec.children) walk(ch, ce.childCellType);
    };
    for (auto& spec : cfg.physicalCells) walk(spec, spec.cellType);
    for (auto& kv : fullCellList)
      for (int32_t l = 1; l <= kv.second.len(); l++)
        for (Cell* c : kv.second.at(l)) {
          c->id = (int32_t)physicalCells.size();
          physicalCells.push_back(c);
        }
    for (auto& vk : vcSchedulers) {
      for (auto& cc : vk.second->nonPinnedFullCellList)
        for (int32_t l = 1; l <= cc.second.len(); l++)
          for (Cell* c : cc.second.at(l)) {
            c->id = (int32_t)virtualCells.size();
            virtualCells.push_back(c);
          }
      for (auto& pc : vk.second->pinnedCells)
        for (int32_t l = 1; l <= pc.second.len(); l++)
          for (Cell* c : pc.second.at(l)) {
            c->id = (int32_t)virtualCells.size();
            virtualCells.push_back(c);
          }
    }
  }

  initCellNums();
  initPinnedCells(P.pinnedPhysicalList);
  initBadNodes();
}

HivedAlgorithm::HivedAlgorithm(const std::string& specText) { parseConfig(specText); }

HivedAlgorithm::~HivedAlgorithm() {
  // a group reached through a stale cell pointer can be deleted a second time by the reference's own flow (Go's GC does
  // not care; Filtering-phase fuzz seed 1059): every group object is freed once
  std::unordered_set<Group*> owned;
  for (auto& kv : affinityGroups) owned.insert(kv.second);
  for (Group* g : deletedGroups) owned.insert(g);
  for (Group* g : owned) delete g;
  for (auto& kv : pods) delete kv.second;
}

Pod* HivedAlgorithm::getPod(int32_t id, int32_t node) {
  auto it = pods.find(id);
  if (it != pods.end()) {
    if (node >= 0) it->second->node = node;
    return it->second;
  }
  Pod* p = new Pod();
  p->id = id;
  p->node = node;
  pods[id] = p;
  return p;
}

// hived_algorithm.go:365-409
void HivedAlgorithm::initCellNums() {
  for (auto& vk : vcFreeCellNum) {
    const std::string& vc = vk.first;
    vcDoomedBadCells[vc];
    for (auto& ck : vk.second) {
      const std::string& chain = ck.first;
      vcDoomedBadCells[vc][chain];
      allVCFreeCellNum[chain];
      for (auto& lk : ck.second) allVCFreeCellNum[chain][lk.first] += lk.second;
    }
  }
  for (auto& ck : allVCFreeCellNum) {
    const std::string& chain = ck.first;
    auto& chainFreeCellNum = ck.second;
    auto fit = fullCellList.find(chain);
    if (fit == fullCellList.end())
      throw Panic("Illegal initial VC assignment: Chain " + chain + " does not exists in physical cluster", HIVED_ERR_BAD_CONFIG);
    const ChainCellList& ccl = fit->second;
    int32_t top = ccl.len();
    int32_t available = (int32_t)ccl.at(top).size();
    totalLeftCellNum[chain];
    badFreeCells[chain];
    allVCDoomedBadCellNum[chain];
    totalLeftCellNum[chain][top] = available;
    for (int32_t l = top; l >= lowestLevel; l--) {
      int32_t need = chainFreeCellNum.count(l) ? chainFreeCellNum[l] : 0;
      int32_t left = available - need;
      if (left < 0) {
        throw Panic("Illegal initial VC assignment: Insufficient physical cells at chain " + chain + " level " +
                        std::to_string(l) + ": " + std::to_string(need) + " needed, " + std::to_string(available) +
                        " available",
                    HIVED_ERR_BAD_CONFIG);
      }
      if (l > lowestLevel) {
        int32_t childNum = (int32_t)ccl.at(l)[0]->children.size();
        available = left * childNum;
        totalLeftCellNum[chain][l - 1] = totalLeftCellNum[chain][l] * childNum;
      }
    }
  }
}

static void bindCell(Cell* pc, Cell* vc);

// hived_algorithm.go:437-449 (maps iterated in ascending key order)
void HivedAlgorithm::initPinnedCells(const std::map<std::string, std::map<std::string, Cell*>>& pinnedPcl) {
  for (auto& vk : pinnedPcl) {
    for (auto& pk : vk.second) {
      Cell* pinnedPhysical = pk.second;
      allocatePreassignedCell(pinnedPhysical, vk.first, false);
      const ChainCellList& virtualList = vcSchedulers[vk.first]->pinnedCells[pk.first];
      Cell* pinnedVirtual = virtualList.at(virtualList.len())[0];
      bindCell(pinnedPhysical, pinnedVirtual);
    }
  }
}

// hived_algorithm.go:451-464
void HivedAlgorithm::initBadNodes() {
  for (auto& kv : fullCellList) {
    const ChainCellList& ccl = kv.second;
    for (Cell* c : ccl.at(ccl.len()))
      for (const std::string& n : std::vector<std::string>(c->nodes)) setBadNode(n);
  }
}

// hived_algorithm.go:466-481
void HivedAlgorithm::setBadNode(const std::string& nodeName) {
  if (badNodes.count(nodeName)) return;
  badNodes.insert(nodeName);
  for (auto& kv : fullCellList) {
    for (Cell* leafCell : kv.second.at(1)) {
      if (leafCell->nodes[0] == nodeName) setBadCell(leafCell);
    }
  }
}

// hived_algorithm.go:483-498
void HivedAlgorithm::setHealthyNode(const std::string& nodeName) {
  if (!badNodes.count(nodeName)) return;
  badNodes.erase(nodeName);
  for (auto& kv : fullCellList) {
    for (Cell* leafCell : kv.second.at(1)) {
      if (leafCell->nodes[0] == nodeName) setHealthyCell(leafCell);
    }
  }
}

// cell_allocation.go:374-382
static Cell* getUnboundVirtualCell(const CellList& cl) {
  for (Cell* c : cl)
    if (c->physicalCell == nullptr) return c;
  return nullptr;
}

// hived_algorithm.go:500-522
void HivedAlgorithm::setBadCell(Cell* c) {
  if (!c->healthy) return;
  SetHealthiness(c, false);
  if (c->parent != nullptr) setBadCell(c->parent);
  if (inFreeCellList(c)) {
    addBadFreeCell(c);
  } else if (c->virtualCell == nullptr && !c->split) {
    Cell* pvc = c->parent ? c->parent->virtualCell : nullptr;
    if (pvc == nullptr) throw Panic("setBadCell: nil pointer dereference (parent has no virtual cell): " + c->address);
    Cell* vc = getUnboundVirtualCell(pvc->children);
    if (vc == nullptr) throw Panic("setBadCell: nil pointer dereference (no unbound virtual cell): " + c->address);
    SetVirtualCell(c, vc);
    SetPhysicalCell(vc, c);
  }
}

// hived_algorithm.go:524-560
void HivedAlgorithm::setHealthyCell(Cell* c) {
  if (c->healthy) return;
  SetHealthiness(c, true);
  if (inFreeCellList(c)) {
    removeBadFreeCell(c);
  } else if (Cell* vc = c->virtualCell) {
    if (!c->pinned && c->priority < minGuaranteedPriority) {
      SetVirtualCell(c, nullptr);
      SetPhysicalCell(vc, nullptr);
      if (vc->parent == nullptr) {
        listRemove(vcDoomedBadCells[vc->vc][c->chain].mut(c->level), c);
        allVCDoomedBadCellNum[c->chain][c->level]--;
        releasePreassignedCell(c, vc->vc, true);
      }
    }
  }
  if (c->parent == nullptr) return;
  for (Cell* buddy : c->parent->children)
    if (!buddy->healthy) return;
  setHealthyCell(c->parent);
}

static int32_t mapGet(const std::map<int32_t, int32_t>& m, int32_t k) {
  auto it = m.find(k);
  return it == m.end() ? 0 : it->second;
}

// hived_algorithm.go:562-581
void HivedAlgorithm::addBadFreeCell(Cell* c) {
  const std::string& chain = c->chain;
  int32_t level = c->level;
  auto bit = badFreeCells.find(chain);
  if (bit == badFreeCells.end()) throw Panic("assignment to entry in nil map (chain " + chain + " belongs to no VC)");
  bit->second.mut(level).push_back(c);
  if (mapGet(allVCFreeCellNum[chain], level) >
      mapGet(totalLeftCellNum[chain], level) - (int32_t)badFreeCells[chain].at(level).size()) {
    tryBindDoomedBadCell(chain, level);
  }
}

// hived_algorithm.go:583-600
void HivedAlgorithm::removeBadFreeCell(Cell* c) {
  const std::string& chain = c->chain;
  int32_t level = c->level;
  listRemove(badFreeCells[chain].mut(level), c);
  tryUnbindDoomedBadCell(chain, level);
}

// hived_algorithm.go:602-628 — VC loop in ascending VC name (Go map order canonicalised)
void HivedAlgorithm::tryBindDoomedBadCell(const std::string& c, int32_t l) {
  for (auto& vk : vcFreeCellNum) {
    const std::string& vcName = vk.first;
    auto cit = vk.second.find(c);
    if (cit == vk.second.end()) continue;
    auto& vcFreeNumC = cit->second;
    while (mapGet(vcFreeNumC, l) > mapGet(totalLeftCellNum[c], l) - (int32_t)badFreeCells[c].at(l).size()) {
      if (badFreeCells[c].at(l).empty()) throw Panic("tryBindDoomedBadCell: index out of range");
      Cell* pc = badFreeCells[c].at(l)[0];
      static const ChainCellList emptyCcl;
      auto& npc = vcSchedulers[vcName]->nonPinnedPreassignedCells;
      auto npit = npc.find(c);
      Cell* vc = getUnboundVirtualCell((npit == npc.end() ? emptyCcl : npit->second).at(l));
      if (vc == nullptr) throw Panic("tryBindDoomedBadCell: nil virtual cell");
      SetVirtualCell(pc, vc);
      SetPhysicalCell(vc, pc);
      vcDoomedBadCells[vcName][c].mut(l).push_back(pc);
      allVCDoomedBadCellNum[c][l]++;
      allocatePreassignedCell(pc, vcName, true);
    }
  }
}

// hived_algorithm.go:630-653
void HivedAlgorithm::tryUnbindDoomedBadCell(const std::string& c, int32_t l) {
  for (auto& vk : vcFreeCellNum) {
    const std::string& vcName = vk.first;
    auto cit = vk.second.find(c);
    if (cit == vk.second.end()) continue;
    auto& vcFreeNumC = cit->second;
    while (!vcDoomedBadCells[vcName][c].at(l).empty() &&
           mapGet(vcFreeNumC, l) < mapGet(totalLeftCellNum[c], l) - (int32_t)badFreeCells[c].at(l).size()) {
      Cell* pc = vcDoomedBadCells[vcName][c].at(l)[0];
      // (pc.GetVirtualCell().GetAddress() / .SetPhysicalCell(nil) on a nil virtual cell: a Go panic in the reference)
      if (pc->virtualCell == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (doomed bad cell is not bound)");
      SetPhysicalCell(pc->virtualCell, nullptr);
      SetVirtualCell(pc, nullptr);
      listRemove(vcDoomedBadCells[vcName][c].mut(l), pc);
      allVCDoomedBadCellNum[c][l]--;
      releasePreassignedCell(pc, vcName, true);
    }
  }
}

// ------------------------------------------------------------------------------------------
// utils.go
// ------------------------------------------------------------------------------------------

// utils.go:381-391
bool inFreeCellList(Cell* c) {
  while (true) {
    if (c->virtualCell != nullptr || c->split) return false;
    if (c->parent == nullptr || c->parent->split) return true;
    c = c->parent;
  }
}

// utils.go:407-415
static bool allChildrenSameState(Cell* c, int32_t s) {
  for (Cell* child : c->children)
    if (child->state != s) return false;
  return true;
}

// utils.go:397-405
void setCellState(Cell* c, int32_t s) {
  SetState(c, s);
  if (c->parent != nullptr) {
    Cell* parent = c->parent;
    if (s == cellUsed || allChildrenSameState(parent, s)) setCellState(parent, s);
  }
}

// utils.go:175-200
static std::set<std::string> collectBadOrNonSuggestedNodes(const Placement& placement,
                                                          const std::unordered_set<std::string>& suggestedNodes,
                                                          bool ignoreSuggestedNodes) {
  std::set<std::string> out;
  for (auto& lk : placement.m)
    for (auto& pod : lk.second)
      for (Cell* leafCell : pod) {
        if (leafCell == nullptr) continue;
        if (!leafCell->healthy || (!ignoreSuggestedNodes && !suggestedNodes.count(leafCell->nodes[0])))
          out.insert(leafCell->nodes[0]);
      }
  return out;
}

// utils.go:202-235; victims flattened to (pod, node) and de-duplicated by pod
static void collectPreemptionVictims(const Placement& placement, std::vector<std::pair<Pod*, int32_t>>& victimPods,
                                     std::vector<Group*>& overlappingPreemptorGroups) {
  std::set<Pod*> seen;
  std::set<Group*> seenG;
  for (auto& lk : placement.m)
    for (auto& pod : lk.second)
      for (Cell* leafCell : pod) {
        if (leafCell == nullptr) continue;
        int32_t state = leafCell->state;
        if (state == cellUsed || state == cellReserving) {
          // pLeafCell.GetUsingGroup().allocatedPods (utils.go:217) on a nil group: a Go panic
          if (leafCell->usingGroup == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (a used cell without a using group)");
          for (auto& pk : leafCell->usingGroup->allocatedPods)
            for (Pod* v : pk.second)
              if (v != nullptr && !seen.count(v)) {
                seen.insert(v);
                victimPods.push_back({v, v->node});
              }
        }
        if (state == cellReserving || state == cellReserved) {
          Group* g = leafCell->reservingOrReservedGroup;
          if (!seenG.count(g)) {
            seenG.insert(g);
            overlappingPreemptorGroups.push_back(g);
          }
        }
      }
  std::sort(victimPods.begin(), victimPods.end(),
            [](const std::pair<Pod*, int32_t>& a, const std::pair<Pod*, int32_t>& b) { return a.first->id < b.first->id; });
  // overlapping preemptors: the reference iterates a Go set; canonical order = group id
  // (a nil entry — a Reserving / Reserved cell without a reserving group, utils.go:229 adds it all the same — sorts first)
  std::sort(overlappingPreemptorGroups.begin(), overlappingPreemptorGroups.end(),
            [](Group* a, Group* b) { return (a ? a->id : -1) < (b ? b->id : -1); });
}

// utils.go:267-283
static Cell* retrieveVirtualCell(const Placement& physicalPlacement, const Placement& virtualPlacement, Cell* pLeafCell) {
  for (auto& lk : physicalPlacement.m)
    for (size_t podIndex = 0; podIndex < lk.second.size(); podIndex++)
      for (size_t leafCellIndex = 0; leafCellIndex < lk.second[podIndex].size(); leafCellIndex++) {
        Cell* leafCell = lk.second[podIndex][leafCellIndex];
        if (leafCell != nullptr && CellEqual(leafCell, pLeafCell))
          return virtualPlacement.m.at(lk.first)[podIndex][leafCellIndex];
      }
  return nullptr;
}

// utils.go:286-296
static int32_t getNewPodIndex(const std::vector<Pod*>& pods) {
  for (size_t i = 0; i < pods.size(); i++)
    if (pods[i] == nullptr) return (int32_t)i;
  return -1;
}

// utils.go:306-316
static bool allPodsReleased(const std::map<int32_t, std::vector<Pod*>>& allocatedPods) {
  for (auto& kv : allocatedPods)
    for (Pod* p : kv.second)
      if (p != nullptr) return false;
  return true;
}

// utils.go:347-378 — LINEAR scan over every leaf of the chain, string compares
Cell* HivedAlgorithm::findPhysicalLeafCellInChain(const std::string& chain, const std::string& node, int32_t leafCellIndex) {
  auto it = fullCellList.find(chain);
  if (it == fullCellList.end()) return nullptr;
  for (Cell* cc : it->second.at(1)) {
    bool success = false;
    for (const std::string& n : cc->nodes)
      if (n == node) {
        success = true;
        break;
      }
    if (success) {
      if (leafCellIndex < 0) return cc;
      for (int32_t g : cc->leafCellIndices)
        if (g == leafCellIndex) return cc;
    }
  }
  return nullptr;
}

// utils.go:318-345 (other chains in ascending name order)
Cell* HivedAlgorithm::findPhysicalLeafCell(const std::string& chain, const std::string& node, int32_t leafCellIndex) {
  Cell* g = findPhysicalLeafCellInChain(chain, node, leafCellIndex);
  if (g == nullptr) {
    for (auto& kv : fullCellList) {
      if (kv.first != chain) {
        g = findPhysicalLeafCellInChain(kv.first, node, leafCellIndex);
        if (g != nullptr) return g;
      }
    }
    return nullptr;
  }
  return g;
}

// utils.go:38-79 + 108-171.  Builds api.PodBindInfo (members ascending by leaf number).
void HivedAlgorithm::generatePodScheduleResult(ScheduleResult& r, int32_t currentLeafCellNum, int32_t currentPodIndex,
                                               Group* group, const std::string& groupName) {
  if (r.physical.nil) {
    r.kind = HIVED_KIND_WAIT;
    return;
  }
  if (!r.victims.empty()) {
    r.kind = HIVED_KIND_PREEMPT;
    return;
  }
  r.kind = HIVED_KIND_BIND;
  r.wait = WaitReason();
  PodBindInfo& info = r.bindInfo;
  info = PodBindInfo();
  std::string chain;
  for (auto& lk : r.physical.m) {  // generateAffinityGroupBindInfo utils.go:108-171
    int32_t podLeafCellNum = lk.first;
    const std::vector<CellList>& podPhysicalPlacements = lk.second;
    std::vector<PodPlacementInfo> mbi(podPhysicalPlacements.size());
    for (size_t podIndex = 0; podIndex < podPhysicalPlacements.size(); podIndex++) {
      mbi[podIndex].physicalLeafCellIndices.assign(podLeafCellNum, 0);
      mbi[podIndex].preassignedCellTypes.assign(podLeafCellNum, "");
      for (int32_t leafCellIndex = 0; leafCellIndex < podLeafCellNum; leafCellIndex++) {
        Cell* pLeafCell = podPhysicalPlacements[podIndex][leafCellIndex];
        if (pLeafCell == nullptr) {
          if (group == nullptr || group->state == groupPreempting)
            throw Panic("The first pod in group " + groupName + " was allocated invalid resource");
          // retrieveMissingPodPlacement (utils.go:250-265) reads the bind-info annotations of the group's other
          // pods.  The annotations live on the pod objects, i.e. above the ABI: the cell is reported as nil
          // (include/hived.h: hived_result_t.incomplete) and the shim completes the pod's placement.
          info.incomplete = true;
          mbi[podIndex].physicalLeafCellIndices[leafCellIndex] = HIVED_NIL_CELL;
          mbi[podIndex].preassignedCellTypes[leafCellIndex] = kNilCellType;
          mbi[podIndex].hasNil = true;
          continue;
        }
        if (mbi[podIndex].physicalNode.empty()) mbi[podIndex].physicalNode = pLeafCell->nodes[0];
        mbi[podIndex].physicalLeafCellIndices[leafCellIndex] = pLeafCell->leafCellIndices[0];
        if (!r.virtual_.nil) {
          Cell* vLeafCell = r.virtual_.m.at(podLeafCellNum)[podIndex][leafCellIndex];
          mbi[podIndex].preassignedCellTypes[leafCellIndex] = cellTypes[vLeafCell->chain][vLeafCell->preassignedCell->level];
        } else {
          mbi[podIndex].preassignedCellTypes[leafCellIndex] = "";
        }
      }
    }
    if (podLeafCellNum == currentLeafCellNum) {
      info.node = mbi[currentPodIndex].physicalNode;
      info.leafCellIsolation = mbi[currentPodIndex].physicalLeafCellIndices;
      Cell* pLeafCell = r.physical.m.at(currentLeafCellNum)[currentPodIndex][0];
      if (pLeafCell != nullptr) chain = pLeafCell->chain;
    }
    info.affinityGroupBindInfo.push_back(mbi);
  }
  info.cellChain = chain;
  r.chain = chain;
}

// ------------------------------------------------------------------------------------------
// cell_allocation.go
// ------------------------------------------------------------------------------------------

// cell_allocation.go:384-397
static void bindCell(Cell* pc, Cell* vc) {
  while (vc->physicalCell == nullptr) {
    SetVirtualCell(pc, vc);
    SetPhysicalCell(vc, pc);
    if (vc->parent == nullptr) break;
    vc = vc->parent;
    pc = pc->parent;
  }
}

// cell_allocation.go:399-420
static void unbindCell(Cell* c) {
  Cell* boundVirtual = c->virtualCell;
  while (!boundVirtual->physicalCell->pinned) {
    Cell* boundPhysical = boundVirtual->physicalCell;
    SetPhysicalCell(boundVirtual, nullptr);
    SetVirtualCell(boundPhysical, nullptr);
    if (boundVirtual->parent == nullptr) return;
    for (Cell* cc : boundVirtual->parent->children)
      if (cc->physicalCell != nullptr) return;
    boundVirtual = boundVirtual->parent;
  }
}

// cell_allocation.go:422-441
void setCellPriority(Cell* c, int32_t p) {
  int32_t originalPriority = c->priority;
  c->priority = p;
  if (Cell* parent = c->parent) {
    if (p > parent->priority) {
      setCellPriority(parent, p);
    } else if (originalPriority == parent->priority && p < originalPriority) {
      int32_t maxBuddyPriority = freePriority;
      for (Cell* buddy : parent->children)
        if (buddy->priority > maxBuddyPriority) maxBuddyPriority = buddy->priority;
      setCellPriority(parent, maxBuddyPriority);
    }
  }
}

// cell_allocation.go:443-454
static void updateUsedLeafCellNumAtPriority(Cell* c, int32_t p, bool increase) {
  while (c != nullptr) {
    IncreaseUsedLeafCellNumAtPriority(c, p, increase ? 1 : -1);
    c = c->parent;
  }
}

// cell_allocation.go:348-372
static Cell* getLowestPriorityVirtualCell(const CellList& cl, int32_t p) {
  int32_t lowestPriority = maxGuaranteedPriority;
  Cell* lowestPriorityCell = nullptr;
  for (Cell* vc : cl) {
    int32_t priority = vc->priority;
    if (priority == freePriority) {
      if (vc->physicalCell == nullptr) return vc;
      continue;
    } else if (priority < p && priority < lowestPriority) {
      lowestPriority = priority;
      lowestPriorityCell = vc;
    }
  }
  return lowestPriorityCell;
}

// cell_allocation.go:317-346
static Cell* mapPhysicalCellToVirtual(Cell* c, const ChainCellList& vccl, int32_t preassignedLevel, int32_t p) {
  if (c->virtualCell != nullptr) return c->virtualCell;
  if (c->level == preassignedLevel) return getLowestPriorityVirtualCell(vccl.at(preassignedLevel), p);
  if (c->parent == nullptr) return nullptr;
  Cell* parentVirtual = mapPhysicalCellToVirtual(c->parent, vccl, preassignedLevel, p);
  if (parentVirtual == nullptr) return nullptr;
  return getLowestPriorityVirtualCell(parentVirtual->children, p);
}

// cell_allocation.go:199-243
bool HivedAlgorithm::getUsablePhysicalCells(const CellList& candidates, int32_t numNeeded,
                                            const std::unordered_set<std::string>& suggestedNodes,
                                            bool ignoreSuggestedNodes, CellList& usableCandidates) {
  usableCandidates.clear();
  stats.free_cells_scanned += (int64_t)candidates.size();
  for (Cell* c : candidates) {
    if (c->virtualCell != nullptr) continue;
    if (c->nodes.size() == 1 && !c->healthy) continue;
    if (!ignoreSuggestedNodes) {
      bool allNonSuggested = true;
      for (const std::string& n : c->nodes)
        if (suggestedNodes.count(n)) {
          allNonSuggested = false;
          break;
        }
      if (allNonSuggested) continue;
    }
    usableCandidates.push_back(c);
  }
  if ((int32_t)usableCandidates.size() < numNeeded) return false;  // nil
  std::stable_sort(usableCandidates.begin(), usableCandidates.end(), [](Cell* a, Cell* b) {
    return usedAt(a, opportunisticPriority) < usedAt(b, opportunisticPriority);
  });
  return true;
}

// cell_allocation.go:245-315
bool HivedAlgorithm::mapVirtualCellsToPhysical(const std::vector<BindingPathVertex*>& cells, const CellList& candidatesIn,
                                               const std::unordered_set<std::string>& suggestedNodes,
                                               bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings,
                                               bool returnPicked, CellList& pickedCells) {
  CellList candidates;
  if (!getUsablePhysicalCells(candidatesIn, (int32_t)cells.size(), suggestedNodes, ignoreSuggestedNodes, candidates))
    return false;
  // a nil result also arises when there are zero usable candidates and zero cells; the reference
  // never calls with zero cells.
  int32_t cellIndex = 0;
  int32_t candidateIndex = 0;
  std::vector<int32_t> pickedCandidateIndices(cells.size(), 0);
  std::set<int32_t> pickedIndexSet;
  while (cellIndex >= 0) {
    for (candidateIndex = pickedCandidateIndices[cellIndex]; candidateIndex < (int32_t)candidates.size(); candidateIndex++) {
      if (pickedIndexSet.count(candidateIndex)) continue;
      Cell* candidate = candidates[candidateIndex];
      bool picked = false;
      if (candidate->level == lowestLevel) {
        picked = true;
        bindings[cells[cellIndex]->cell->address] = candidate;
      } else {
        CellList ignored;
        picked = mapVirtualCellsToPhysical(cells[cellIndex]->childrenToBind, candidate->children, suggestedNodes,
                                           ignoreSuggestedNodes, bindings, false, ignored);
      }
      if (picked) {
        pickedCandidateIndices[cellIndex] = candidateIndex;
        pickedIndexSet.insert(candidateIndex);
        if (cellIndex == (int32_t)cells.size() - 1) {
          if (!returnPicked) return true;
          pickedCells.clear();
          for (int32_t index : pickedCandidateIndices) pickedCells.push_back(candidates[index]);
          return true;
        }
        break;
      }
    }
    if (candidateIndex == (int32_t)candidates.size()) {
      cellIndex--;
      if (cellIndex >= 0) {
        pickedIndexSet.erase(pickedCandidateIndices[cellIndex]);
        pickedCandidateIndices[cellIndex]++;
      }
    } else {
      cellIndex++;
    }
  }
  return false;
}

// cell_allocation.go:34-80
bool HivedAlgorithm::buddyAlloc(BindingPathVertex* cell, ChainCellList& freeList, int32_t currentLevel,
                                const std::unordered_set<std::string>& suggestedNodes, bool ignoreSuggestedNodes,
                                std::map<std::string, Cell*>& bindings) {
  if (currentLevel == cell->cell->level) {
    CellList pickedCells;
    bool ok = mapVirtualCellsToPhysical({cell}, freeList.at(currentLevel), suggestedNodes, ignoreSuggestedNodes,
                                        bindings, true, pickedCells);
    if (ok) {
      for (Cell* c : pickedCells) listRemove(freeList.mut(currentLevel), c);
      return true;
    }
    return false;
  }
  CellList freeCells;
  if (!getUsablePhysicalCells(freeList.at(currentLevel), 1, suggestedNodes, ignoreSuggestedNodes, freeCells)) return false;
  for (Cell* c : freeCells) {
    CellList& lower = freeList.mut(currentLevel - 1);
    lower.insert(lower.end(), c->children.begin(), c->children.end());
    if (buddyAlloc(cell, freeList, currentLevel - 1, suggestedNodes, ignoreSuggestedNodes, bindings)) {
      listRemove(freeList.mut(currentLevel), c);
      return true;
    } else {
      freeList.mut(currentLevel - 1).clear();  // = nil
    }
  }
  return false;
}

// cell_allocation.go:82-150
bool HivedAlgorithm::safeRelaxedBuddyAlloc(BindingPathVertex* cell, ChainCellList& freeList,
                                           std::map<int32_t, int32_t>& freeCellNum, int32_t currentLevel,
                                           const std::unordered_set<std::string>& suggestedNodes,
                                           bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings) {
  Cell* splittableCell = nullptr;
  std::map<int32_t, int32_t> splittableNum;
  int32_t top = freeList.len();
  for (int32_t i = top; i > currentLevel; i--) {
    splittableNum[i] = (int32_t)freeList.at(i).size() - mapGet(freeCellNum, i);
    if (i < top && splittableCell != nullptr)
      splittableNum[i] += splittableNum[i + 1] * (int32_t)splittableCell->children.size();
    if (splittableCell == nullptr && !freeList.at(i).empty()) {
      splittableCell = freeList.at(i)[0];
    } else if (splittableCell != nullptr) {
      splittableCell = splittableCell->children[0];
    }
    if (splittableNum[i] < 0)
      throw Panic("VC Safety Broken: level " + std::to_string(i) + " cell is unsplittable, splittableNum=" +
                  std::to_string(splittableNum[i]));
  }
  for (int32_t l = currentLevel + 1; l <= top; l++) {
    int32_t cellNum = (int32_t)freeList.at(l).size();
    if (cellNum > splittableNum[l]) cellNum = splittableNum[l];
    if (cellNum > 0) {
      CellList splitList;
      for (int32_t i = 0; i < cellNum; i++) {
        Cell* first = freeList.at(l)[0];
        splitList.push_back(first);
        listRemove(freeList.mut(l), first);
      }
      splittableNum[l] -= cellNum;
      for (int32_t sl = l; sl > currentLevel; sl--) {
        CellList splitChildrenList;
        for (Cell* sc : splitList)
          splitChildrenList.insert(splitChildrenList.end(), sc->children.begin(), sc->children.end());
        splitList = splitChildrenList;
      }
      CellList& cur = freeList.mut(currentLevel);
      splitList.insert(splitList.end(), cur.begin(), cur.end());  // prepend
      cur = splitList;
      CellList pickedCells;
      bool ok = mapVirtualCellsToPhysical({cell}, freeList.at(currentLevel), suggestedNodes, ignoreSuggestedNodes,
                                          bindings, true, pickedCells);
      if (ok) {
        for (Cell* c : pickedCells) listRemove(freeList.mut(currentLevel), c);
        return true;
      }
    }
  }
  return false;
}

// cell_allocation.go:152-161
static int32_t getLowestFreeCellLevel(const ChainCellList& freeList, int32_t l) {
  for (; l <= freeList.len(); l++)
    if (!freeList.at(l).empty()) return l;
  throw Panic("VC Safety Broken: free cell not found even split to the highest level " + std::to_string(l - 1));
}

// cell_allocation.go:163-197
bool HivedAlgorithm::mapVirtualPlacementToPhysical(std::vector<BindingPathVertex*>& preassignedCells,
                                                   std::vector<std::vector<BindingPathVertex*>>& nonPreassignedCells,
                                                   ChainCellList& freeList, std::map<int32_t, int32_t>& freeCellNum,
                                                   const std::unordered_set<std::string>& suggestedNodes,
                                                   bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings) {
  for (BindingPathVertex* c : preassignedCells) {
    if (!buddyAlloc(c, freeList, getLowestFreeCellLevel(freeList, c->cell->level), suggestedNodes,
                    ignoreSuggestedNodes, bindings)) {
      if (!safeRelaxedBuddyAlloc(c, freeList, freeCellNum, c->cell->level, suggestedNodes, ignoreSuggestedNodes, bindings))
        return false;
    } else {
      freeCellNum[c->cell->level]--;
    }
  }
  for (auto& cells : nonPreassignedCells) {
    CellList ignored;
    bool ok = mapVirtualCellsToPhysical(cells, cells[0]->cell->parent->physicalCell->children, suggestedNodes,
                                        ignoreSuggestedNodes, bindings, false, ignored);
    if (!ok) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// types.go
// ------------------------------------------------------------------------------------------

BindingPathVertex* HivedAlgorithm::newVertex(Cell* c) {
  auto up = std::make_unique<BindingPathVertex>();
  up->cell = c;
  BindingPathVertex* v = up.get();
  vertexStore_.push_back(std::move(up));
  return v;
}

// types.go:282-340
void HivedAlgorithm::toBindingPaths(const Placement& p, const std::vector<int32_t>& leafCellNums,
                                    std::map<std::string, Cell*>& bindings,
                                    std::vector<BindingPathVertex*>& preassignedCells,
                                    std::vector<std::vector<BindingPathVertex*>>& nonPreassignedCells) {
  std::map<std::string, BindingPathVertex*> allBindingPathVertices;
  for (int32_t podLeafCellNum : leafCellNums) {
    const std::vector<CellList>& podPlacements = p.m.at(podLeafCellNum);
    for (const CellList& podPlacement : podPlacements) {
      for (Cell* leafCell : podPlacement) {
        if (Cell* pLeafCell = leafCell->physicalCell) {
          bindings[leafCell->address] = pLeafCell;
          continue;
        }
        std::vector<Cell*> bindingPath;
        for (Cell* c = leafCell; c != nullptr; c = c->parent) {
          if (c->physicalCell != nullptr || allBindingPathVertices.count(c->address)) break;
          bindingPath.push_back(c);
        }
        Cell* pathRoot = bindingPath.back();
        BindingPathVertex* n = newVertex(pathRoot);
        allBindingPathVertices[pathRoot->address] = n;
        if (Cell* parent = pathRoot->parent; parent == nullptr) {
          preassignedCells.push_back(n);
        } else if (parent->physicalCell != nullptr) {
          bool buddyExist = false;
          for (auto& group : nonPreassignedCells) {
            if (CellEqual(parent, group[0]->cell->parent)) {
              buddyExist = true;
              group.push_back(n);
              break;
            }
          }
          if (!buddyExist) nonPreassignedCells.push_back({n});
        } else {
          BindingPathVertex* parentNode = allBindingPathVertices.at(pathRoot->parent->address);
          parentNode->childrenToBind.push_back(n);
        }
        for (int i = (int)bindingPath.size() - 2; i >= 0; i--) {
          Cell* c = bindingPath[i];
          BindingPathVertex* nn = newVertex(c);
          BindingPathVertex* parentNode = allBindingPathVertices.at(c->parent->address);
          parentNode->childrenToBind.push_back(nn);
          allBindingPathVertices[c->address] = nn;
        }
      }
    }
  }
}

// types.go:260-280
static Placement toPhysicalPlacement(const Placement& p, const std::map<std::string, Cell*>& bindings,
                                     const std::vector<int32_t>& leafCellNums) {
  Placement physicalPlacement;
  physicalPlacement.nil = false;
  for (int32_t podLeafCellNum : leafCellNums) {
    const std::vector<CellList>& podPlacements = p.m.at(podLeafCellNum);
    std::vector<CellList>& out = physicalPlacement.m[podLeafCellNum];
    out.resize(podPlacements.size());
    for (size_t i = 0; i < podPlacements.size(); i++) {
      out[i].resize(podPlacements[i].size());
      for (size_t j = 0; j < podPlacements[i].size(); j++) {
        auto it = bindings.find(podPlacements[i][j]->address);
        out[i][j] = it == bindings.end() ? nullptr : it->second;
      }
    }
  }
  return physicalPlacement;
}

// types.go:150-183
static Group* newAlgoAffinityGroup(const PodSchedulingSpec& s, int32_t state) {
  Group* group = new Group();
  for (auto& m : s.members) group->totalPodNums[m.second] += m.first;
  group->name = s.groupName;
  group->id = s.groupId;
  group->vc = s.virtualCluster;
  group->lazyPreemptionEnable = s.lazyPreemptionEnable;
  group->priority = s.priority;
  group->state = state;
  group->physicalPlacement.nil = false;
  group->virtualPlacement.nil = false;
  for (auto& kv : group->totalPodNums) {
    int32_t leafCellNum = kv.first, podNum = kv.second;
    group->physicalPlacement.m[leafCellNum].assign(podNum, CellList(leafCellNum, nullptr));
    group->virtualPlacement.m[leafCellNum].assign(podNum, CellList(leafCellNum, nullptr));
    group->allocatedPods[leafCellNum].assign(podNum, nullptr);
  }
  return group;
}

// ------------------------------------------------------------------------------------------
// topology_aware_scheduler.go
// ------------------------------------------------------------------------------------------

// topology_aware_scheduler.go:181-189
static Cell* ancestorNoHigherThanNode(Cell* c) {
  if (c->atOrHigherThanNode || c->parent == nullptr) return c;
  return ancestorNoHigherThanNode(c->parent);
}

// topology_aware_scheduler.go:51-63 + newClusterView :158-179
TopologyAwareScheduler* HivedAlgorithm::newTopologyAwareScheduler(const ChainCellList& ccl,
                                                                  const std::map<int32_t, int32_t>& levelLeafCellNum,
                                                                  bool crossPriorityPack) {
  auto up = std::make_unique<TopologyAwareScheduler>();
  TopologyAwareScheduler* t = up.get();
  schedStore_.push_back(std::move(up));
  t->levelLeafCellNum = levelLeafCellNum;
  t->crossPriorityPack = crossPriorityPack;
  int32_t l;
  for (l = 1; l <= ccl.len(); l++) {
    if (ccl.at(l)[0]->atOrHigherThanNode) break;
  }
  for (; l >= lowestLevel; l--) {
    for (Cell* c : ccl.at(l)) {
      Cell* a = ancestorNoHigherThanNode(c);
      bool contains = false;  // cv.containsCell :191-198
      for (Node* n : t->cv)
        if (CellEqual(a, n->c)) {
          contains = true;
          break;
        }
      if (!contains) {
        auto nu = std::make_unique<Node>();
        nu->c = c;
        t->cv.push_back(nu.get());
        nodeStore_.push_back(std::move(nu));
      }
    }
  }
  return t;
}

// topology_aware_scheduler.go:138-154
static void updateUsedLeafCellNumForPriority(Node* n, int32_t p, bool crossPriorityPack) {
  n->usedLeafCellNumSamePriority = usedAt(n->c, p);
  n->usedLeafCellNumHigherPriority = 0;
  n->freeLeafCellNumAtPriority = n->c->totalLeafCellNum;
  for (auto& kv : n->c->used) {
    int32_t priority = kv.first, num = kv.second;
    if (crossPriorityPack) {
      if (priority != p) n->usedLeafCellNumSamePriority += num;
    } else if (priority > p) {
      n->usedLeafCellNumHigherPriority += num;
    }
    if (priority >= p) n->freeLeafCellNumAtPriority -= num;
  }
}

// topology_aware_scheduler.go:242-265
static void nodeHealthyAndInSuggested(Node* n, const std::unordered_set<std::string>& suggestedNodes,
                                      bool ignoreSuggestedNodes) {
  if (n->c->physical) {
    n->healthy = n->c->healthy;
    n->suggested = ignoreSuggestedNodes || suggestedNodes.count(n->c->nodes[0]) != 0;
    n->nodeAddressCell = n->c;
    return;
  }
  if (Cell* pn = n->c->physicalCell) {
    n->healthy = pn->healthy;
    n->suggested = ignoreSuggestedNodes || suggestedNodes.count(pn->nodes[0]) != 0;
    n->nodeAddressCell = pn;
    return;
  }
  n->healthy = true;
  n->suggested = true;
  n->nodeAddressCell = nullptr;
}

// topology_aware_scheduler.go:205-224
static bool nodeLess(const Node* a, const Node* b) {
  if (a->healthy != b->healthy) return a->healthy;
  if (a->suggested != b->suggested) return a->suggested;
  if (a->usedLeafCellNumSamePriority > b->usedLeafCellNumSamePriority) return true;
  if (a->usedLeafCellNumSamePriority < b->usedLeafCellNumSamePriority) return false;
  if (a->usedLeafCellNumHigherPriority < b->usedLeafCellNumHigherPriority) return true;
  return false;
}

// topology_aware_scheduler.go:267-306
static bool findNodesForPods(std::vector<Node*>& cv, const std::vector<int32_t>& leafCellNums,
                             std::vector<int32_t>& pickedNodeIndices, WaitReason& failedReason) {
  std::stable_sort(cv.begin(), cv.end(), nodeLess);  // sort.Stable, in place: order persists across calls
  pickedNodeIndices.assign(leafCellNums.size(), 0);
  size_t podIndex = 0;
  int32_t pickedLeafCellNum = 0;
  for (size_t nodeIndex = 0; nodeIndex < cv.size();) {
    Node* n = cv[nodeIndex];
    if (n->freeLeafCellNumAtPriority - pickedLeafCellNum >= leafCellNums[podIndex]) {
      if (!n->healthy) {
        failedReason.code = HIVED_WAIT_BAD_NODE;
        failedReason.cell = n->nodeAddressCell;
        return false;
      }
      if (!n->suggested) {
        failedReason.code = HIVED_WAIT_NON_SUGGESTED_NODE;
        failedReason.cell = n->nodeAddressCell;
        return false;
      }
      pickedNodeIndices[podIndex] = (int32_t)nodeIndex;
      pickedLeafCellNum += leafCellNums[podIndex];
      podIndex++;
      if (podIndex == leafCellNums.size()) {
        failedReason = WaitReason();
        return true;
      }
    } else {
      pickedLeafCellNum = 0;
      nodeIndex++;
    }
  }
  failedReason.code = HIVED_WAIT_INSUFFICIENT;
  failedReason.cell = nullptr;
  return false;
}

// topology_aware_scheduler.go:443-462
Cell* findLCA(Cell* lower, Cell* higher) {
  while (lower->level < higher->level) {
    if (lower->parent == nullptr) return nullptr;
    lower = lower->parent;
  }
  if (CellEqual(lower, higher)) return lower;
  while (!CellEqual(lower->parent, higher->parent)) {
    if (lower->parent == nullptr || higher->parent == nullptr) return nullptr;
    lower = lower->parent;
    higher = higher->parent;
  }
  return lower->parent;
}

// topology_aware_scheduler.go:464-476
static void getLeafCellsFromNode(Cell* c, int32_t p, CellList& freeLeafCells, CellList& preemptibleLeafCells) {
  if (c->level > 1) {
    for (Cell* cc : c->children) getLeafCellsFromNode(cc, p, freeLeafCells, preemptibleLeafCells);
  } else if (c->priority == freePriority) {
    freeLeafCells.push_back(c);
  } else if (c->priority < p) {
    preemptibleLeafCells.push_back(c);
  }
}

// topology_aware_scheduler.go:389-399
static int32_t getOptimalAffinity(int32_t leafCellNum, const std::map<int32_t, int32_t>& levelLeafCellNum) {
  for (int32_t l = 1; l <= (int32_t)levelLeafCellNum.size(); l++)
    if (mapGet(levelLeafCellNum, l) >= leafCellNum) return l;
  throw Panic("Assert Failure: pod allocated a node but exceeds the capacity of the current chain");
}

// topology_aware_scheduler.go:425-441
static void removePickedLeafCells(CellList& leafCells, const std::vector<int32_t>& indices) {
  for (size_t i = 0; i < indices.size(); i++) {
    int32_t index = indices[i];
    int32_t offset = (int32_t)i;
    if (i < indices.size() - 1) {
      int32_t nextIndex = indices[i + 1];
      std::copy(leafCells.begin() + index + 1, leafCells.begin() + nextIndex, leafCells.begin() + index - offset);
    } else {
      std::copy(leafCells.begin() + index + 1, leafCells.end(), leafCells.begin() + index - offset);
    }
  }
  leafCells.resize(leafCells.size() - indices.size());
}

// topology_aware_scheduler.go:308-387.  availableLeafCells: nil-ness carried by `haveList`.
static CellList findLeafCellsInNode(Cell* n, int32_t leafCellNum, int32_t p, CellList& availableLeafCells,
                                    bool& haveList, const std::map<int32_t, int32_t>& levelLeafCellNum) {
  std::vector<int32_t> currentLeafCellIndices(leafCellNum, 0);
  CellList currentAffinity(leafCellNum, nullptr);
  CellList bestAffinityLeafCells(leafCellNum, nullptr);
  std::vector<int32_t> bestAffinityLeafCellIndices(leafCellNum, 0);
  int32_t bestAffinity = highestLevel;
  int32_t optimalAffinity = getOptimalAffinity(leafCellNum, levelLeafCellNum);

  if (!haveList) {
    availableLeafCells.clear();
    CellList preemptibleLeafCells;
    getLeafCellsFromNode(n, p, availableLeafCells, preemptibleLeafCells);
    availableLeafCells.insert(availableLeafCells.end(), preemptibleLeafCells.begin(), preemptibleLeafCells.end());
    haveList = true;
  }
  int32_t availableLeafCellIndex = 0;
  int32_t searchLeafCellIndex = 0;
  while (true) {
    while (availableLeafCellIndex < (int32_t)availableLeafCells.size()) {
      Cell* leafCell = availableLeafCells[availableLeafCellIndex];
      currentLeafCellIndices[searchLeafCellIndex] = availableLeafCellIndex;
      if (searchLeafCellIndex == 0) {
        currentAffinity[searchLeafCellIndex] = leafCell;
      } else {
        currentAffinity[searchLeafCellIndex] = findLCA(leafCell, currentAffinity[searchLeafCellIndex - 1]);
        if ((currentAffinity[searchLeafCellIndex] == nullptr && bestAffinity < highestLevel) ||
            (currentAffinity[searchLeafCellIndex] != nullptr && currentAffinity[searchLeafCellIndex]->level > bestAffinity)) {
          availableLeafCellIndex++;
          continue;
        }
      }
      if (searchLeafCellIndex == leafCellNum - 1) {
        // checkCurrentLeafCells :401-423
        Cell* last = currentAffinity[leafCellNum - 1];
        if (last == nullptr) throw Panic("findLeafCellsInNode: nil pointer dereference (no common ancestor)");
        int32_t affinity = last->level;
        bool foundOptimalAffinity = false;
        if (affinity < bestAffinity) {
          bestAffinityLeafCellIndices = currentLeafCellIndices;
          for (int32_t i = 0; i < leafCellNum; i++) bestAffinityLeafCells[i] = availableLeafCells[currentLeafCellIndices[i]];
          bestAffinity = affinity;
          foundOptimalAffinity = affinity == optimalAffinity;
        }
        if (foundOptimalAffinity) {
          removePickedLeafCells(availableLeafCells, bestAffinityLeafCellIndices);
          return bestAffinityLeafCells;
        }
      } else {
        searchLeafCellIndex++;
      }
      availableLeafCellIndex++;
    }
    searchLeafCellIndex--;
    if (searchLeafCellIndex < 0) {
      if (bestAffinity == highestLevel)
        throw Panic("Assert Failure: failed to allocate " + std::to_string(leafCellNum) + " leaf cells in picked node " + n->address);
      removePickedLeafCells(availableLeafCells, bestAffinityLeafCellIndices);
      return bestAffinityLeafCells;
    }
    availableLeafCellIndex = currentLeafCellIndices[searchLeafCellIndex] + 1;
  }
}

// topology_aware_scheduler.go:65-116 (+ updateClusterView :231-240)
void HivedAlgorithm::tasSchedule(TopologyAwareScheduler* t, const std::map<int32_t, int32_t>& podLeafCellNumbers,
                                 int32_t p, const std::unordered_set<std::string>& suggestedNodes,
                                 bool ignoreSuggestedNodes, Placement& podPlacements, WaitReason& failedReason) {
  std::vector<int32_t> sortedPodLeafCellNumbers;
  for (auto& kv : podLeafCellNumbers)
    for (int32_t i = 0; i < kv.second; i++) sortedPodLeafCellNumbers.push_back(kv.first);
  std::sort(sortedPodLeafCellNumbers.begin(), sortedPodLeafCellNumbers.end());

  int32_t priority = opportunisticPriority;
  auto updateClusterView = [&](int32_t pr) {
    stats.view_nodes_scanned += (int64_t)t->cv.size();
    for (Node* n : t->cv) {
      updateUsedLeafCellNumForPriority(n, pr, t->crossPriorityPack);
      nodeHealthyAndInSuggested(n, suggestedNodes, ignoreSuggestedNodes);
    }
  };
  updateClusterView(priority);
  std::vector<int32_t> selectedNodeIndices;
  bool ok = findNodesForPods(t->cv, sortedPodLeafCellNumbers, selectedNodeIndices, failedReason);
  if (!ok && p > opportunisticPriority) {
    priority = p;
    updateClusterView(priority);
    ok = findNodesForPods(t->cv, sortedPodLeafCellNumbers, selectedNodeIndices, failedReason);
  }
  if (!ok) {
    podPlacements = Placement();
    return;
  }
  stats.pods_placed += (int64_t)sortedPodLeafCellNumbers.size();
  CellList selectedNodes(sortedPodLeafCellNumbers.size());
  for (size_t i = 0; i < selectedNodeIndices.size(); i++) selectedNodes[i] = t->cv[selectedNodeIndices[i]]->c;
  std::map<Cell*, std::pair<bool, CellList>> nodeAvailableLeafCells;
  podPlacements = Placement();
  podPlacements.nil = false;
  for (size_t podIndex = 0; podIndex < sortedPodLeafCellNumbers.size(); podIndex++) {
    int32_t leafCellNumber = sortedPodLeafCellNumbers[podIndex];
    Cell* n = selectedNodes[podIndex];
    auto& entry = nodeAvailableLeafCells[n];
    CellList selectedLeafCells = findLeafCellsInNode(n, leafCellNumber, priority, entry.second, entry.first, t->levelLeafCellNum);
    podPlacements.m[leafCellNumber].push_back(selectedLeafCells);
  }
  failedReason = WaitReason();
}

// intra_vc_scheduler.go:92-117
void HivedAlgorithm::intraVCSchedule(IntraVCScheduler* s, SchedulingRequest& sr, Placement& placement, WaitReason& failedReason) {
  TopologyAwareScheduler* scheduler = nullptr;
  if (!sr.pinnedCellId.empty()) {
    auto it = s->pinnedCellSchedulers.find(sr.pinnedCellId);
    if (it != s->pinnedCellSchedulers.end()) scheduler = it->second;
  } else {
    auto it = s->nonPinnedCellSchedulers.find(sr.chain);
    if (it != s->nonPinnedCellSchedulers.end()) scheduler = it->second;
  }
  placement = Placement();
  failedReason = WaitReason();
  if (scheduler != nullptr) {
    tasSchedule(scheduler, sr.affinityGroupPodNums, sr.priority, *sr.suggestedNodes, sr.ignoreSuggestedNodes, placement, failedReason);
  } else {
    failedReason.code = HIVED_WAIT_NO_SCHEDULER;
  }
  if (placement.nil) {
    failedReason.code |= HIVED_WAIT_SCOPE_VC;
    return;
  }
  failedReason = WaitReason();
}

// ------------------------------------------------------------------------------------------
// hived_algorithm.go — Schedule and the group state machine
// ------------------------------------------------------------------------------------------

// hived_algorithm.go:180-224
ScheduleResult HivedAlgorithm::Schedule(const PodSchedulingSpec& s, int32_t podId,
                                        const std::vector<std::string>& suggestedNodes, bool preemptingPhase) {
  std::unordered_set<std::string> suggestedNodeSet;
  for (const std::string& n : suggestedNodes) suggestedNodeSet.insert(n);
  ScheduleResult r;
  vertexStore_.clear();
  auto git = affinityGroups.find(s.groupName);
  if (git != affinityGroups.end()) {
    schedulePodFromExistingGroup(git->second, s, suggestedNodeSet, preemptingPhase, podId, r);
  }
  if (affinityGroups.find(s.groupName) == affinityGroups.end()) {
    schedulePodFromNewGroup(s, suggestedNodeSet, preemptingPhase, podId, r);
  }
  git = affinityGroups.find(s.groupName);
  generatePodScheduleResult(r, s.leafCellNumber, r.podIndex, git == affinityGroups.end() ? nullptr : git->second, s.groupName);
  return r;
}

// hived_algorithm.go:229-245
void HivedAlgorithm::DeleteUnallocatedPod(const std::string& groupName, int32_t podId) {
  auto git = affinityGroups.find(groupName);
  if (git != affinityGroups.end() && git->second->state == groupPreempting) {
    Group* g = git->second;
    g->preemptingPods.erase(podId);
    if (g->preemptingPods.empty()) deletePreemptingAffinityGroup(g);
  }
}

// hived_algorithm.go:247-270
void HivedAlgorithm::AddAllocatedPod(const PodSchedulingSpec& s, const PodBindInfo& info, int32_t podId,
                                     int32_t nodeId, int32_t podIndexFromInfo) {
  int32_t podIndex = 0;
  auto git = affinityGroups.find(s.groupName);
  if (git != affinityGroups.end()) {
    Group* g = git->second;
    if (g->state == groupPreempting) allocatePreemptingAffinityGroup(g);
    podIndex = podIndexFromInfo;  // getAllocatedPodIndex(info, s.LeafCellNumber), utils.go:291-304
    if (podIndex == -1) return;
  } else {
    createAllocatedAffinityGroup(s, info);
  }
  Group* g = affinityGroups.at(s.groupName);
  auto& slots = g->allocatedPods[s.leafCellNumber];
  if (podIndex < 0 || podIndex >= (int32_t)slots.size()) throw Panic("AddAllocatedPod: index out of range");
  slots[podIndex] = getPod(podId, nodeId);
}

// hived_algorithm.go:272-296
// returns the id of the pod that occupied the cleared slot (-1: nothing cleared) — hived_result_t.pod_index of the event
int32_t HivedAlgorithm::DeleteAllocatedPod(const std::string& groupName, int32_t leafCellNumber, int32_t podIndex) {
  lastRemovedPod = -1;
  auto git = affinityGroups.find(groupName);
  if (git == affinityGroups.end()) return -1;
  Group* g = git->second;
  if (podIndex == -1) return -1;
  auto& slots = g->allocatedPods[leafCellNumber];
  if (podIndex < 0 || podIndex >= (int32_t)slots.size()) throw Panic("DeleteAllocatedPod: index out of range");
  const int32_t occupant = slots[podIndex] ? slots[podIndex]->id : -1;
  slots[podIndex] = nullptr;
  lastRemovedPod = occupant;  // (also when the group's deletion below panics)
  if (allPodsReleased(g->allocatedPods)) deleteAllocatedAffinityGroup(g);
  return occupant;
}

// hived_algorithm.go:655-712
void HivedAlgorithm::schedulePodFromExistingGroup(Group* g, const PodSchedulingSpec& s,
                                                  const std::unordered_set<std::string>& suggestedNodes,
                                                  bool preemptingPhase, int32_t podId, ScheduleResult& r) {
  std::set<std::string> badOrNonSuggestedNodes =
      collectBadOrNonSuggestedNodes(g->physicalPlacement, suggestedNodes, g->ignoreK8sSuggestedNodes);
  if (g->state == groupAllocated) {
    r.physical = g->physicalPlacement;
    r.virtual_ = g->virtualPlacement;
    r.podIndex = getNewPodIndex(g->allocatedPods[s.leafCellNumber]);
    if (r.podIndex == -1)
      throw BadRequest(HIVED_ERR_TOO_MANY_PODS, "Requesting more pods than the configured number for " +
                                                    std::to_string(s.leafCellNumber) + " leaf cells (" +
                                                    std::to_string(mapGet(g->totalPodNums, s.leafCellNumber)) +
                                                    " pods) in affinity group " + s.groupName);
  } else {  // groupPreempting (a BeingPreempted group keeps state Allocated semantics? no: see below)
    // NB the reference's else-branch also receives groupBeingPreempted; it is restated literally.
    if (preemptingPhase && !badOrNonSuggestedNodes.empty()) {
      deletePreemptingAffinityGroup(g);
    } else {
      r.physical = g->physicalPlacement;
      r.virtual_ = g->virtualPlacement;
      std::vector<Group*> ignored;
      collectPreemptionVictims(r.physical, r.victims, ignored);
      g->preemptingPods[podId] = getPod(podId, -1);
    }
  }
}

// hived_algorithm.go:714-752
void HivedAlgorithm::schedulePodFromNewGroup(const PodSchedulingSpec& s,
                                             const std::unordered_set<std::string>& suggestedNodes,
                                             bool preemptingPhase, int32_t podId, ScheduleResult& r) {
  scheduleNewAffinityGroup(s, suggestedNodes, r.physical, r.virtual_, r.wait);
  if (r.physical.nil) {
    r.virtual_ = Placement();
    r.victims.clear();
    return;
  }
  std::vector<Group*> overlappingPreemptors;
  r.victims.clear();
  collectPreemptionVictims(r.physical, r.victims, overlappingPreemptors);
  if (preemptingPhase) {
    for (Group* preemptor : overlappingPreemptors) {
      // deletePreemptingAffinityGroup(nil) reads g.physicalLeafCellPlacement (:1116): a Go panic
      if (preemptor == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (cancelling the preemption of a nil group)");
      deletePreemptingAffinityGroup(preemptor);
    }
    if (!r.victims.empty()) createPreemptingAffinityGroup(s, r.physical, r.virtual_, podId);
  }
}

// hived_algorithm.go:754-796
void HivedAlgorithm::scheduleNewAffinityGroup(const PodSchedulingSpec& s,
                                              const std::unordered_set<std::string>& suggestedNodes,
                                              Placement& phys, Placement& virt, WaitReason& failedReason) {
  SchedulingRequest sr;
  sr.vc = s.virtualCluster;
  sr.pinnedCellId = s.pinnedCellId;
  sr.priority = s.priority;
  sr.affinityGroupName = s.groupName;
  sr.suggestedNodes = &suggestedNodes;
  sr.ignoreSuggestedNodes = s.ignoreK8sSuggestedNodes;
  for (auto& m : s.members) sr.affinityGroupPodNums[m.second] += m.first;
  validateSchedulingRequest(sr);
  if (!sr.pinnedCellId.empty()) {
    handleSchedulingRequest(sr, phys, virt, failedReason);
  } else if (!s.leafCellType.empty()) {
    if (!cellChains.count(s.leafCellType))
      throw BadRequest(HIVED_ERR_LEAF_TYPE_NOT_IN_CLUSTER,
                       "Pod requesting leaf cell type " + s.leafCellType + " which the whole cluster does not have");
    scheduleAffinityGroupForLeafCellType(sr, s.leafCellType, true, phys, virt, failedReason);
  } else {
    scheduleAffinityGroupForAnyLeafCellType(sr, phys, virt, failedReason);
  }
}

// hived_algorithm.go:798-829
void HivedAlgorithm::scheduleAffinityGroupForLeafCellType(SchedulingRequest& sr, const std::string& leafCellType,
                                                          bool typeSpecified, Placement& phys, Placement& virt,
                                                          WaitReason& failedReason) {
  bool vcHasType = false;
  failedReason = WaitReason();
  phys = Placement();
  virt = Placement();
  for (const std::string& chain : cellChains[leafCellType]) {
    if (sr.priority < minGuaranteedPriority || vcSchedulers[sr.vc]->nonPinnedPreassignedCells.count(chain)) {
      vcHasType = true;
      sr.chain = chain;
      handleSchedulingRequest(sr, phys, virt, failedReason);
      if (!phys.nil) {
        failedReason = WaitReason();
        return;
      }
    }
  }
  if (typeSpecified && sr.priority >= minGuaranteedPriority && !vcHasType)
    throw BadRequest(HIVED_ERR_LEAF_TYPE_NOT_IN_VC,
                     "Pod requesting leaf cell type " + leafCellType + " which VC " + sr.vc + " does not have");
  phys = Placement();
  virt = Placement();
}

// hived_algorithm.go:831-853 (leaf types in ascending name order)
void HivedAlgorithm::scheduleAffinityGroupForAnyLeafCellType(SchedulingRequest& sr, Placement& phys, Placement& virt,
                                                             WaitReason& failedReasonOut) {
  WaitReason failedReason;
  for (auto& kv : cellChains) {
    WaitReason typeFailedReason;
    scheduleAffinityGroupForLeafCellType(sr, kv.first, false, phys, virt, typeFailedReason);
    if (!phys.nil) {
      failedReasonOut = WaitReason();
      return;
    }
    if (!typeFailedReason.empty()) failedReason = typeFailedReason;
  }
  phys = Placement();
  virt = Placement();
  failedReasonOut = failedReason;
}

// hived_algorithm.go:855-870
void HivedAlgorithm::validateSchedulingRequest(const SchedulingRequest& sr) {
  auto it = vcSchedulers.find(sr.vc);
  if (it == vcSchedulers.end()) throw BadRequest(HIVED_ERR_UNKNOWN_VC, "VC " + sr.vc + " does not exists!");
  if (!sr.pinnedCellId.empty()) {
    if (!it->second->pinnedCells.count(sr.pinnedCellId))
      throw BadRequest(HIVED_ERR_UNKNOWN_PINNED_CELL, "VC " + sr.vc + " does not have pinned cell " + sr.pinnedCellId);
    if (sr.priority == opportunisticPriority)
      throw BadRequest(HIVED_ERR_OPPORTUNISTIC_PINNED, "opportunistic pod not supported to use pinned cell " + sr.pinnedCellId);
  }
}

// hived_algorithm.go:872-896
void HivedAlgorithm::handleSchedulingRequest(SchedulingRequest& sr, Placement& phys, Placement& virt, WaitReason& failedReason) {
  if (sr.priority >= minGuaranteedPriority) {
    scheduleGuaranteedAffinityGroup(sr, phys, virt, failedReason);
  } else {
    virt = Placement();
    scheduleOpportunisticAffinityGroup(sr, phys, failedReason);
  }
  if (phys.nil) {
    virt = Placement();
    return;
  }
  failedReason = WaitReason();
}

// hived_algorithm.go:898-942
void HivedAlgorithm::scheduleGuaranteedAffinityGroup(SchedulingRequest& sr, Placement& phys, Placement& virtualPlacement,
                                                     WaitReason& failedReason) {
  intraVCSchedule(vcSchedulers[sr.vc], sr, virtualPlacement, failedReason);
  if (virtualPlacement.nil) {
    phys = Placement();
    return;
  }
  std::map<std::string, Cell*> bindings;
  std::vector<int32_t> leafCellNums;
  for (auto& kv : sr.affinityGroupPodNums) leafCellNums.push_back(kv.first);
  std::sort(leafCellNums.begin(), leafCellNums.end());
  std::map<std::string, Placement> lazyPreemptedGroups = tryLazyPreempt(virtualPlacement, leafCellNums, sr.affinityGroupName);
  std::vector<BindingPathVertex*> preassignedCells;
  std::vector<std::vector<BindingPathVertex*>> nonPreassignedCells;
  toBindingPaths(virtualPlacement, leafCellNums, bindings, preassignedCells, nonPreassignedCells);
  std::map<int32_t, int32_t> freeCellNumCopy;
  if (auto it = allVCFreeCellNum.find(sr.chain); it != allVCFreeCellNum.end()) freeCellNumCopy = it->second;
  ChainCellList freeListCopy;  // sr.chain is "" for pinned-cell requests: nil list in the reference
  if (auto it = freeCellList.find(sr.chain); it != freeCellList.end()) freeListCopy = shallowCopy(it->second);
  if (mapVirtualPlacementToPhysical(preassignedCells, nonPreassignedCells, freeListCopy, freeCellNumCopy,
                                    *sr.suggestedNodes, sr.ignoreSuggestedNodes, bindings)) {
    phys = toPhysicalPlacement(virtualPlacement, bindings, leafCellNums);
    failedReason = WaitReason();
    return;
  }
  for (auto& kv : lazyPreemptedGroups) revertLazyPreempt(affinityGroups.at(kv.first), kv.second);
  phys = Placement();
  virtualPlacement = Placement();
  failedReason.code = HIVED_WAIT_MAPPING;
  failedReason.cell = nullptr;
}

// hived_algorithm.go:944-965
std::map<std::string, Placement> HivedAlgorithm::tryLazyPreempt(const Placement& p, const std::vector<int32_t>& leafCellNums,
                                                                const std::string& groupName) {
  std::map<std::string, Placement> preemptedGroups;
  for (int32_t podLeafCellNum : leafCellNums) {
    for (const CellList& pod : p.m.at(podLeafCellNum)) {
      for (Cell* leafCell : pod) {
        if (Cell* pLeafCell = leafCell->physicalCell) {
          if (pLeafCell->state == cellUsed && pLeafCell->usingGroup->lazyPreemptionEnable) {
            Group* victim = pLeafCell->usingGroup;
            preemptedGroups[victim->name] = lazyPreemptAffinityGroup(victim, groupName);
          }
        }
      }
    }
  }
  return preemptedGroups;
}

// hived_algorithm.go:967-979
void HivedAlgorithm::scheduleOpportunisticAffinityGroup(SchedulingRequest& sr, Placement& placement, WaitReason& failedReason) {
  tasSchedule(opportunisticSchedulers.at(sr.chain), sr.affinityGroupPodNums, opportunisticPriority, *sr.suggestedNodes,
              sr.ignoreSuggestedNodes, placement, failedReason);
  if (placement.nil) {
    failedReason.code |= HIVED_WAIT_SCOPE_PHYSICAL;
    return;
  }
  failedReason = WaitReason();
}

// hived_algorithm.go:981-1041
void HivedAlgorithm::createAllocatedAffinityGroup(const PodSchedulingSpec& s, const PodBindInfo& info) {
  Group* newGroup = newAlgoAffinityGroup(s, groupAllocated);
  std::unique_ptr<Group> guard(newGroup);
  bool shouldLazyPreempt = false;
  for (auto& gms : info.affinityGroupBindInfo) {
    int32_t leafCellNumber = (int32_t)gms[0].physicalLeafCellIndices.size();
    for (int32_t podIndex = 0; podIndex < (int32_t)gms.size(); podIndex++) {
      const std::string& node = gms[podIndex].physicalNode;
      for (int32_t leafCellIndex = 0; leafCellIndex < (int32_t)gms[podIndex].physicalLeafCellIndices.size(); leafCellIndex++) {
        Cell* pLeafCell = nullptr;
        Cell* vLeafCell = nullptr;
        int lazyPreempt = 0;
        findAllocatedLeafCell(leafCellIndex, gms[podIndex].physicalLeafCellIndices, gms[podIndex], info.cellChain, node,
                              shouldLazyPreempt, s, newGroup, pLeafCell, vLeafCell, lazyPreempt);
        if (pLeafCell == nullptr) continue;
        auto& physSlots = newGroup->physicalPlacement.m[leafCellNumber];
        if (podIndex >= (int32_t)physSlots.size() || leafCellIndex >= (int32_t)physSlots[podIndex].size())
          throw Panic("createAllocatedAffinityGroup: index out of range");
        physSlots[podIndex][leafCellIndex] = pLeafCell;
        if (lazyPreempt == 0) {
          newGroup->virtualPlacement = Placement();  // = nil
        } else if (vLeafCell != nullptr) {
          newGroup->virtualPlacement.m[leafCellNumber][podIndex][leafCellIndex] = vLeafCell;
          if (inFreeCellList(pLeafCell) && vLeafCell->preassignedCell->priority > freePriority)
            lazyPreemptCell(vLeafCell->preassignedCell, newGroup->name);
        } else {
          shouldLazyPreempt = shouldLazyPreempt || (lazyPreempt == 2);
        }
        bool safetyOk = allocateLeafCell(pLeafCell, vLeafCell, s.priority, newGroup->vc);
        pLeafCell->usingGroup = newGroup;  // AddUsingGroup cell.go:218-225
        setCellState(pLeafCell, cellUsed);
        if (!safetyOk) shouldLazyPreempt = true;
      }
    }
  }
  if (shouldLazyPreempt) lazyPreemptAffinityGroup(newGroup, newGroup->name);
  affinityGroups[s.groupName] = guard.release();
}

// hived_algorithm.go:1043-1070
void HivedAlgorithm::deleteAllocatedAffinityGroup(Group* g) {
  for (auto& lk : g->physicalPlacement.m)
    for (auto& podPlacement : lk.second)
      for (Cell* pLeafCell : podPlacement) {
        if (pLeafCell == nullptr) continue;
        pLeafCell->usingGroup = nullptr;  // DeleteUsingGroup cell.go:227-233
        if (pLeafCell->state == cellUsed) {
          releaseLeafCell(pLeafCell, g->vc);
          setCellState(pLeafCell, cellFree);
        } else {
          setCellState(pLeafCell, cellReserved);
        }
      }
  affinityGroups.erase(g->name);
  deletedGroups.push_back(g);  // Go's GC keeps a deleted group alive while a cell still points to it (a stale usingGroup is READ by the reference)
}

// hived_algorithm.go:1072-1112
void HivedAlgorithm::createPreemptingAffinityGroup(const PodSchedulingSpec& s, const Placement& physicalPlacement,
                                                   const Placement& virtualPlacement, int32_t podId) {
  Group* newGroup = newAlgoAffinityGroup(s, groupPreempting);
  newGroup->physicalPlacement = physicalPlacement;
  newGroup->virtualPlacement = virtualPlacement;
  for (auto& lk : physicalPlacement.m) {
    int32_t leafCellNum = lk.first;
    for (size_t podIndex = 0; podIndex < lk.second.size(); podIndex++) {
      for (size_t leafCellIndex = 0; leafCellIndex < lk.second[podIndex].size(); leafCellIndex++) {
        Cell* pLeafCell = lk.second[podIndex][leafCellIndex];
        Cell* vLeafCell = virtualPlacement.m.at(leafCellNum)[podIndex][leafCellIndex];
        if (pLeafCell->state == cellUsed) {
          Group* usingGroup = pLeafCell->usingGroup;
          releaseLeafCell(pLeafCell, usingGroup->vc);
          usingGroup->state = groupBeingPreempted;
        }
        allocateLeafCell(pLeafCell, vLeafCell, s.priority, newGroup->vc);
        pLeafCell->reservingOrReservedGroup = newGroup;
        if (pLeafCell->state == cellUsed) {
          setCellState(pLeafCell, cellReserving);
        } else {
          setCellState(pLeafCell, cellReserved);
        }
      }
    }
  }
  newGroup->preemptingPods[podId] = getPod(podId, -1);
  affinityGroups[s.groupName] = newGroup;
}

// hived_algorithm.go:1114-1145
void HivedAlgorithm::deletePreemptingAffinityGroup(Group* g) {
  for (auto& lk : g->physicalPlacement.m)
    for (auto& podPlacement : lk.second)
      for (Cell* pLeafCell : podPlacement) {
        releaseLeafCell(pLeafCell, g->vc);
        pLeafCell->reservingOrReservedGroup = nullptr;
        if (pLeafCell->state == cellReserving) {
          setCellState(pLeafCell, cellUsed);
          Group* beingPreemptedGroup = pLeafCell->usingGroup;
          Cell* beingPreemptedVLeafCell = nullptr;
          if (!beingPreemptedGroup->virtualPlacement.nil)
            beingPreemptedVLeafCell = retrieveVirtualCell(beingPreemptedGroup->physicalPlacement,
                                                          beingPreemptedGroup->virtualPlacement, pLeafCell);
          allocateLeafCell(pLeafCell, beingPreemptedVLeafCell, beingPreemptedGroup->priority, beingPreemptedGroup->vc);
        } else {
          setCellState(pLeafCell, cellFree);
        }
      }
  affinityGroups.erase(g->name);
  deletedGroups.push_back(g);  // Go's GC keeps a deleted group alive while a cell still points to it (a stale usingGroup is READ by the reference)
}

// hived_algorithm.go:1147-1163
void HivedAlgorithm::allocatePreemptingAffinityGroup(Group* g) {
  for (auto& lk : g->physicalPlacement.m)
    for (auto& podPlacement : lk.second)
      for (Cell* pLeafCell : podPlacement) {
        pLeafCell->reservingOrReservedGroup = nullptr;
        pLeafCell->usingGroup = g;
        setCellState(pLeafCell, cellUsed);
      }
  g->state = groupAllocated;
  g->preemptingPods.clear();
}

// hived_algorithm.go:1165-1191
Placement HivedAlgorithm::lazyPreemptAffinityGroup(Group* victim, const std::string& preemptor) {
  (void)preemptor;
  for (auto& lk : victim->virtualPlacement.m)
    for (auto& podVirtualPlacement : lk.second)
      for (Cell* vLeafCell : podVirtualPlacement) {
        if (vLeafCell != nullptr) {
          Cell* pLeafCell = vLeafCell->physicalCell;
          // (a virtual leaf that health churn left unbound: releaseLeafCell reads pLeafCell.virtualCell through a nil
          // *PhysicalCell in the reference, :1331 — a Go panic)
          if (pLeafCell == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (lazy preemption of an unbound virtual leaf cell)");
          releaseLeafCell(pLeafCell, victim->vc);
          allocateLeafCell(pLeafCell, nullptr, opportunisticPriority, victim->vc);
        }
      }
  Placement originalVirtualPlacement = victim->virtualPlacement;
  victim->virtualPlacement = Placement();
  victim->lazyPreempted = true;
  return originalVirtualPlacement;
}

// hived_algorithm.go:1193-1201
void HivedAlgorithm::lazyPreemptCell(Cell* c, const std::string& preemptor) {
  if (c->level == lowestLevel && c->state == cellUsed) lazyPreemptAffinityGroup(c->physicalCell->usingGroup, preemptor);
  for (Cell* child : c->children) lazyPreemptCell(child, preemptor);
}

// hived_algorithm.go:1203-1222
void HivedAlgorithm::revertLazyPreempt(Group* g, const Placement& virtualPlacement) {
  for (auto& lk : g->physicalPlacement.m)
    for (size_t podIndex = 0; podIndex < lk.second.size(); podIndex++)
      for (size_t leafCellIndex = 0; leafCellIndex < lk.second[podIndex].size(); leafCellIndex++) {
        Cell* pLeafCell = lk.second[podIndex][leafCellIndex];
        if (pLeafCell == nullptr) continue;
        Cell* vLeafCell = virtualPlacement.m.at(lk.first)[podIndex][leafCellIndex];
        releaseLeafCell(pLeafCell, g->vc);
        allocateLeafCell(pLeafCell, vLeafCell, g->priority, g->vc);
      }
  g->virtualPlacement = virtualPlacement;
  g->lazyPreempted = false;
}

// hived_algorithm.go:1224-1290
void HivedAlgorithm::findAllocatedLeafCell(int32_t index, const std::vector<int32_t>& physicalLeafCellIndices,
                                           const PodPlacementInfo& pp, const std::string& chain, const std::string& node,
                                           bool lazyPreempted, const PodSchedulingSpec& s, Group* group, Cell*& pLeafCell,
                                           Cell*& vLeafCell, int& lazyPreempt) {
  int32_t priority = s.priority;
  int32_t physicalLeafCellIndex = physicalLeafCellIndices[index];
  pLeafCell = findPhysicalLeafCell(chain, node, physicalLeafCellIndex);
  vLeafCell = nullptr;
  if (pLeafCell == nullptr) {
    lazyPreempt = 1;
    return;
  }
  if (pp.preassignedNil) {
    lazyPreempt = 2;
    return;
  }
  if (!group->virtualPlacement.nil && !lazyPreempted) {
    const std::string& preassignedType = pp.preassignedCellTypes[index];
    if (!preassignedType.empty()) {
      int32_t preassignedLevel = 0;
      bool typeFound = false;
      for (auto& lt : cellTypes[pLeafCell->chain])
        if (lt.second == preassignedType) {
          preassignedLevel = lt.first;
          typeFound = true;
        }
      if (typeFound) {
        auto vit = vcSchedulers.find(s.virtualCluster);
        if (vit != vcSchedulers.end()) {
          const ChainCellList* vccl = nullptr;
          if (!s.pinnedCellId.empty()) {
            auto pit = vit->second->pinnedCells.find(s.pinnedCellId);
            if (pit != vit->second->pinnedCells.end()) vccl = &pit->second;
          } else {
            auto cit = vit->second->nonPinnedPreassignedCells.find(pLeafCell->chain);
            if (cit != vit->second->nonPinnedPreassignedCells.end()) vccl = &cit->second;
          }
          if (vccl != nullptr) vLeafCell = mapPhysicalCellToVirtual(pLeafCell, *vccl, preassignedLevel, priority);
        }
      }
      lazyPreempt = vLeafCell == nullptr ? 2 : 1;
      return;
    }
    lazyPreempt = 0;  // nil: opportunistic group without virtual placement
    return;
  }
  lazyPreempt = 1;
}

// hived_algorithm.go:1292-1323
bool HivedAlgorithm::allocateLeafCell(Cell* pLeafCell, Cell* vLeafCell, int32_t p, const std::string& vcn) {
  bool safetyOk = true;
  stats.leaves_committed++;
  if (vLeafCell != nullptr) {
    setCellPriority(vLeafCell, p);
    updateUsedLeafCellNumAtPriority(vLeafCell, p, true);
    setCellPriority(pLeafCell, p);
    updateUsedLeafCellNumAtPriority(pLeafCell, p, true);
    Cell* pac = vLeafCell->preassignedCell;
    bool preassignedNewlyBound = pac->physicalCell == nullptr;
    if (pLeafCell->virtualCell == nullptr) bindCell(pLeafCell, vLeafCell);
    if (preassignedNewlyBound) safetyOk = allocatePreassignedCell(pac->physicalCell, vcn, false);
  } else {
    setCellPriority(pLeafCell, opportunisticPriority);
    updateUsedLeafCellNumAtPriority(pLeafCell, opportunisticPriority, true);
  }
  return safetyOk;
}

// hived_algorithm.go:1325-1352
void HivedAlgorithm::releaseLeafCell(Cell* pLeafCell, const std::string& vcn) {
  stats.leaves_committed++;
  if (Cell* vLeafCell = pLeafCell->virtualCell) {
    updateUsedLeafCellNumAtPriority(vLeafCell, vLeafCell->priority, false);
    setCellPriority(vLeafCell, freePriority);
    Cell* preassignedPhysical = vLeafCell->preassignedCell->physicalCell;
    // (nil when health churn left the preassigned cell unbound above a bound leaf: IsPinned() on it is a nil-pointer
    // dereference in the reference, i.e. a Go panic = platform error)
    if (preassignedPhysical == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (preassigned cell is not bound)");
    if (pLeafCell->healthy) unbindCell(pLeafCell);
    if (!preassignedPhysical->pinned && vLeafCell->preassignedCell->priority < minGuaranteedPriority &&
        !listContains(vcDoomedBadCells[vcn][preassignedPhysical->chain].at(preassignedPhysical->level), preassignedPhysical)) {
      releasePreassignedCell(preassignedPhysical, vcn, false);
    }
  }
  updateUsedLeafCellNumAtPriority(pLeafCell, pLeafCell->priority, false);
  setCellPriority(pLeafCell, freePriority);
}

// hived_algorithm.go:1354-1427
bool HivedAlgorithm::allocatePreassignedCell(Cell* c, const std::string& vcn, bool doomedBad) {
  bool safetyOk = true;
  // allocateLeafCell (:1312-1316) passes preassignedCell.GetPhysicalCell(), which is still nil when the physical leaf was
  // already bound to another virtual cell (bindCell skipped): c.GetChain() on a nil *PhysicalCell is a Go panic
  if (c == nullptr) throw Panic("runtime error: invalid memory address or nil pointer dereference (allocating a nil preassigned cell)");
  const std::string chain = c->chain;
  int32_t level = c->level;
  // h.vcFreeCellNum[vcn][chain][level]-- on a VC that has no counters for the chain writes into a nil map: a Go panic
  if (!vcFreeCellNum[vcn].count(chain)) throw Panic("assignment to entry in nil map (VC " + vcn + " has no cells of chain " + chain + ")");
  vcFreeCellNum[vcn][chain][level]--;
  allVCFreeCellNum[chain][level]--;
  totalLeftCellNum[chain][level]--;
  int32_t splitLevelUpTo = removeCellFromFreeList(c);
  Cell* parent = c->parent;
  for (int32_t l = level + 1; l <= splitLevelUpTo; l++) {
    totalLeftCellNum[chain][l]--;
    if (mapGet(totalLeftCellNum[chain], l) < mapGet(allVCFreeCellNum[chain], l)) safetyOk = false;
    if (!parent->healthy) {
      listRemove(badFreeCells[chain].mut(l), parent);
    } else {
      tryBindDoomedBadCell(chain, l);
    }
    parent = parent->parent;
  }
  if (!c->healthy) {
    allocateBadCell(c);
    if (!doomedBad) tryUnbindDoomedBadCell(chain, level);
  } else {
    tryBindDoomedBadCell(chain, level);
  }
  int32_t numToReduce = (int32_t)c->children.size();
  for (int32_t l = level - 1; l >= lowestLevel; l--) {
    totalLeftCellNum[chain][l] -= numToReduce;
    if (mapGet(totalLeftCellNum[chain], l) < mapGet(allVCFreeCellNum[chain], l)) safetyOk = false;
    if (!doomedBad) tryBindDoomedBadCell(chain, l);
    numToReduce *= (int32_t)fullCellList[chain].at(l)[0]->children.size();
  }
  return safetyOk;
}

// hived_algorithm.go:1429-1447
void HivedAlgorithm::allocateBadCell(Cell* c) {
  if (listContains(badFreeCells[c->chain].at(c->level), c)) listRemove(badFreeCells[c->chain].mut(c->level), c);
  if (c->virtualCell == nullptr) {
    Cell* pvc = c->parent ? c->parent->virtualCell : nullptr;
    if (pvc == nullptr) throw Panic("allocateBadCell: nil pointer dereference: " + c->address);
    Cell* vc = getUnboundVirtualCell(pvc->children);
    if (vc == nullptr) throw Panic("allocateBadCell: nil virtual cell: " + c->address);
    SetVirtualCell(c, vc);
    SetPhysicalCell(vc, c);
  }
  for (Cell* child : c->children)
    if (!child->healthy) allocateBadCell(child);
}

// hived_algorithm.go:1449-1485
void HivedAlgorithm::releasePreassignedCell(Cell* c, const std::string& vcn, bool doomedBad) {
  const std::string chain = c->chain;
  int32_t level = c->level;
  if (!vcFreeCellNum[vcn].count(chain)) throw Panic("assignment to entry in nil map (VC " + vcn + " has no cells of chain " + chain + ")");
  vcFreeCellNum[vcn][chain][level]++;
  allVCFreeCellNum[chain][level]++;
  totalLeftCellNum[chain][level]++;
  int32_t mergeLevelUpTo = addCellToFreeList(c);
  Cell* parent = c->parent;
  for (int32_t l = level + 1; l <= mergeLevelUpTo; l++) {
    totalLeftCellNum[chain][l]++;
    if (!parent->healthy) {
      badFreeCells[chain].mut(l).push_back(parent);
    } else {
      tryUnbindDoomedBadCell(chain, l);
    }
    parent = parent->parent;
  }
  if (!c->healthy) {
    releaseBadCell(c);
    if (!doomedBad) tryBindDoomedBadCell(chain, level);
  } else {
    tryUnbindDoomedBadCell(chain, level);
  }
  int32_t numToAdd = (int32_t)c->children.size();
  for (int32_t l = level - 1; l >= lowestLevel; l--) {
    totalLeftCellNum[chain][l] += numToAdd;
    if (!doomedBad) tryUnbindDoomedBadCell(chain, l);
    numToAdd *= (int32_t)fullCellList[chain].at(l)[0]->children.size();
  }
}

// hived_algorithm.go:1487-1500
void HivedAlgorithm::releaseBadCell(Cell* c) {
  badFreeCells[c->chain].mut(c->level).push_back(c);
  if (Cell* vc = c->virtualCell) {
    SetVirtualCell(c, nullptr);
    SetPhysicalCell(vc, nullptr);
  }
  for (Cell* child : c->children)
    if (!child->healthy) releaseBadCell(child);
}

// hived_algorithm.go:1502-1527
int32_t HivedAlgorithm::removeCellFromFreeList(Cell* c) {
  const std::string chain = c->chain;
  bool terminate = false;
  while (true) {
    int32_t l = c->level;
    Cell* parent = c->parent;
    if (parent != nullptr) {
      if (parent->split) {
        terminate = true;
      } else {
        CellList& fl = freeCellList[chain].mut(l);
        fl.insert(fl.end(), parent->children.begin(), parent->children.end());
        parent->split = true;
      }
    } else {
      terminate = true;
    }
    listRemove(freeCellList[chain].mut(l), c);
    if (terminate) return l;
    c = parent;
  }
}

// hived_algorithm.go:1529-1565
int32_t HivedAlgorithm::addCellToFreeList(Cell* c) {
  const std::string chain = c->chain;
  bool terminate = false;
  while (true) {
    int32_t l = c->level;
    Cell* parent = c->parent;
    if (parent != nullptr) {
      bool allBuddyFree = true;
      for (Cell* buddy : parent->children) {
        if (!CellEqual(buddy, c) && !listContains(freeCellList[chain].at(l), buddy)) {
          allBuddyFree = false;
          break;
        }
      }
      if (!allBuddyFree) {
        terminate = true;
      } else {
        for (Cell* buddy : parent->children)
          if (!CellEqual(buddy, c)) listRemove(freeCellList[chain].mut(l), buddy);
        parent->split = false;
      }
    } else {
      terminate = true;
    }
    if (terminate) {
      freeCellList[chain].mut(l).push_back(c);
      return l;
    }
    c = parent;
  }
}

}  // namespace hived_oracle
