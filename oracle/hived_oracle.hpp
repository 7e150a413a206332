// TEST INFRASTRUCTURE — CPU oracle.  NOT part of the product path.
//
// A faithful, single-threaded, pointer-based C++17 restatement of the reference's scheduling
// algorithm, microsoft/hivedscheduler pkg/algorithm/*.go (Go).  It deliberately keeps the
// reference's data structures (pointer-linked cells, address-string equality, per-call full
// recompute, linear scans, swap-remove lists) so that (a) it can be checked line by line against
// the Go source and (b) it is an honest stand-in for the Go CPU path when timing the baseline
// (the Go toolchain is not available in this image).
//
// Parity is PINNED: tests/test_oracle_golden.py replays the reference's own
// TestHivedAlgorithm scenario (pkg/algorithm/hived_algorithm_test.go:613-632) against this code
// and checks every golden vector (expectedBindInfos :566-592, expectedPreemptInfos :594-602, ...).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  Every function cites the reference file:line it follows.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace hived_oracle {

// constants.go:30-39
constexpr int32_t maxGuaranteedPriority = 1000;
constexpr int32_t minGuaranteedPriority = 0;
constexpr int32_t opportunisticPriority = -1;
constexpr int32_t freePriority = opportunisticPriority - 1;
constexpr int32_t lowestLevel = 1;
constexpr int32_t highestLevel = INT32_MAX;

// constants.go:43-70
enum CellState { cellFree = 0, cellUsed = 1, cellReserving = 2, cellReserved = 3 };
enum GroupState { groupNone = 0, groupAllocated = 1, groupPreempting = 2, groupBeingPreempted = 3 };

// internal.NewBadRequestError panics (pkg/internal/utils.go:316-326) vs plain panics.
struct BadRequest : std::runtime_error {
  int code;
  BadRequest(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct Panic : std::runtime_error {
  int code;
  explicit Panic(const std::string& m, int c = 100) : std::runtime_error(m), code(c) {}
};

struct Group;
struct Cell;
typedef std::vector<Cell*> CellList;  // types.go:55

// cell.go:58-142 (GenericCell + PhysicalCell) and :315-324 (VirtualCell) folded into one record.
struct Cell {
  bool physical = true;
  std::string chain;
  int32_t level = 0;
  std::string address;
  Cell* parent = nullptr;
  CellList children;
  bool atOrHigherThanNode = false;
  bool isNodeLevel = false;
  int32_t priority = freePriority;
  int32_t state = cellFree;
  bool healthy = true;
  int32_t totalLeafCellNum = 0;
  std::map<int32_t, int32_t> used;  // usedLeafCellNumAtPriorities
  // PhysicalCell
  std::vector<std::string> nodes;
  std::vector<int32_t> leafCellIndices;
  Group* usingGroup = nullptr;
  Group* reservingOrReservedGroup = nullptr;
  Cell* virtualCell = nullptr;
  bool split = false;
  bool pinned = false;
  // VirtualCell
  std::string vc;
  std::string pid;
  Cell* preassignedCell = nullptr;
  Cell* physicalCell = nullptr;
  // ABI bookkeeping (not in the reference)
  int32_t id = -1;
  std::string cellType;
};

// types.go:98-130
struct ChainCellList {
  std::map<int32_t, CellList> m;
  int32_t len() const { return (int32_t)m.size(); }
  const CellList& at(int32_t l) const {
    static const CellList empty;
    auto it = m.find(l);
    return it == m.end() ? empty : it->second;
  }
  CellList& mut(int32_t l) { return m[l]; }
  bool has(int32_t l) const { return m.count(l) != 0; }
};

struct Pod {
  int32_t id = -1;
  int32_t node = -1;  // Spec.NodeName (interned) of the binding pod
};

// types.go:228-229; nil-ness of the map is observable in the reference
struct Placement {
  bool nil = true;
  std::map<int32_t, std::vector<CellList>> m;  // leaf cell number -> pods -> leaf cells
};

// types.go:133-148
struct Group {
  std::string name;
  int32_t id = -1;
  std::string vc;
  bool lazyPreemptionEnable = false;
  bool ignoreK8sSuggestedNodes = false;  // never set by the reference (types.go:150-183)
  int32_t priority = 0;
  std::map<int32_t, int32_t> totalPodNums;
  std::map<int32_t, std::vector<Pod*>> allocatedPods;
  std::map<int32_t, Pod*> preemptingPods;  // keyed by pod id (UID)
  Placement physicalPlacement;
  Placement virtualPlacement;
  int32_t state = groupNone;
  bool lazyPreempted = false;  // lazyPreemptionStatus != nil
};

// types.go:43-52
struct SchedulingRequest {
  std::string vc;
  std::string pinnedCellId;
  std::string chain;
  std::string affinityGroupName;
  std::map<int32_t, int32_t> affinityGroupPodNums;
  int32_t priority = 0;
  const std::unordered_set<std::string>* suggestedNodes = nullptr;
  bool ignoreSuggestedNodes = true;
};

// api.PodSchedulingSpec, pkg/api/types.go:78-99
struct PodSchedulingSpec {
  std::string virtualCluster;
  int32_t priority = 0;
  std::string pinnedCellId;
  std::string leafCellType;
  int32_t leafCellNumber = 0;
  bool lazyPreemptionEnable = false;
  bool ignoreK8sSuggestedNodes = true;
  std::string groupName;
  int32_t groupId = -1;
  std::vector<std::pair<int32_t, int32_t>> members;  // (podNumber, leafCellNumber)
};

// api.PodBindInfo, pkg/api/types.go:101-118
struct PodPlacementInfo {
  std::string physicalNode;
  std::vector<int32_t> physicalLeafCellIndices;
  std::vector<std::string> preassignedCellTypes;
  bool preassignedNil = false;
  bool hasNil = false;  // some leaf cell of this pod is nil (left the spec): index HIVED_NIL_CELL, type kNilCellType
};
static const char* const kNilCellType = "\x01nil-cell";
struct PodBindInfo {
  bool incomplete = false;  // hived_result_t.incomplete
  std::string node;
  std::vector<int32_t> leafCellIsolation;
  std::string cellChain;
  std::vector<std::vector<PodPlacementInfo>> affinityGroupBindInfo;  // member -> pods
};

// topology_aware_scheduler.go:118-126
struct Node {
  Cell* c = nullptr;
  int32_t freeLeafCellNumAtPriority = 0;
  int32_t usedLeafCellNumSamePriority = 0;
  int32_t usedLeafCellNumHigherPriority = 0;
  bool healthy = true;
  bool suggested = true;
  Cell* nodeAddressCell = nullptr;  // the physical cell whose address the reference logs
};

// topology_aware_scheduler.go:36-49
struct TopologyAwareScheduler {
  std::vector<Node*> cv;
  std::map<int32_t, int32_t> levelLeafCellNum;
  bool crossPriorityPack = false;
};

// intra_vc_scheduler.go:46-55
struct IntraVCScheduler {
  std::map<std::string, ChainCellList> nonPinnedFullCellList;
  std::map<std::string, ChainCellList> nonPinnedPreassignedCells;
  std::map<std::string, ChainCellList> pinnedCells;
  std::map<std::string, TopologyAwareScheduler*> nonPinnedCellSchedulers;
  std::map<std::string, TopologyAwareScheduler*> pinnedCellSchedulers;
};

// types.go:343-347
struct BindingPathVertex {
  Cell* cell = nullptr;
  std::vector<BindingPathVertex*> childrenToBind;
};

struct WaitReason {
  int32_t code = 0;   // HIVED_WAIT_* | scope
  Cell* cell = nullptr;
  bool empty() const { return code == 0; }
};

struct ScheduleResult {
  int kind = 0;  // 0 wait 1 bind 2 preempt
  WaitReason wait;
  Placement physical;
  Placement virtual_;
  std::vector<std::pair<Pod*, int32_t>> victims;  // (pod, node id)
  int32_t podIndex = 0;
  std::string chain;
  PodBindInfo bindInfo;
};

struct Stats {
  int64_t view_nodes_scanned = 0;
  int64_t leaves_committed = 0;
  int64_t free_cells_scanned = 0;
  int64_t pods_placed = 0;
};

// hived_algorithm.go:40-105
class HivedAlgorithm {
 public:
  explicit HivedAlgorithm(const std::string& specText);
  ~HivedAlgorithm();

  // internal.SchedulerAlgorithm
  ScheduleResult Schedule(const PodSchedulingSpec& s, int32_t podId,
                          const std::vector<std::string>& suggestedNodes, bool preemptingPhase);
  void AddAllocatedPod(const PodSchedulingSpec& s, const PodBindInfo& info, int32_t podId,
                       int32_t nodeId, int32_t podIndexFromInfo);
  int32_t DeleteAllocatedPod(const std::string& groupName, int32_t leafCellNumber, int32_t podIndex);
  int32_t lastRemovedPod = -1;  // occupant of the slot the last DeleteAllocatedPod cleared
  void DeleteUnallocatedPod(const std::string& groupName, int32_t podId);
  void setBadNode(const std::string& nodeName);
  void setHealthyNode(const std::string& nodeName);

  // state (public: the ABI adapter and tests read it)
  std::map<std::string, IntraVCScheduler*> vcSchedulers;
  std::map<std::string, TopologyAwareScheduler*> opportunisticSchedulers;
  std::map<std::string, ChainCellList> fullCellList;
  std::map<std::string, ChainCellList> freeCellList;
  std::unordered_map<std::string, Group*> affinityGroups;
  std::vector<Group*> deletedGroups;  // freed with the algorithm object (see deleteAllocatedAffinityGroup)
  std::map<std::string, std::map<std::string, std::map<int32_t, int32_t>>> vcFreeCellNum;
  std::map<std::string, std::map<int32_t, int32_t>> allVCFreeCellNum;
  std::map<std::string, std::map<int32_t, int32_t>> totalLeftCellNum;
  std::map<std::string, ChainCellList> badFreeCells;
  std::map<std::string, std::map<std::string, ChainCellList>> vcDoomedBadCells;
  std::map<std::string, std::map<int32_t, int32_t>> allVCDoomedBadCellNum;
  std::unordered_set<std::string> badNodes;
  std::map<std::string, std::vector<std::string>> cellChains;           // leaf type -> chains
  std::map<std::string, std::map<int32_t, std::string>> cellTypes;      // chain -> level -> type
  std::map<std::string, std::map<int32_t, int32_t>> leafCellNums;       // chain -> level -> #leaves

  // interning tables for the ABI
  std::vector<std::string> nodeNames, chainNames, vcNames, leafTypeNames, pinnedNames, cellTypeNames;
  std::unordered_map<std::string, int32_t> nodeIds;
  std::vector<Cell*> physicalCells, virtualCells;  // by ABI id
  std::unordered_map<int32_t, Pod*> pods;
  Stats stats;

  Pod* getPod(int32_t id, int32_t node);

 private:
  std::vector<std::unique_ptr<Cell>> cellStore_;
  std::vector<std::unique_ptr<Node>> nodeStore_;
  std::vector<std::unique_ptr<TopologyAwareScheduler>> schedStore_;
  std::vector<std::unique_ptr<IntraVCScheduler>> vcsStore_;
  std::vector<std::unique_ptr<BindingPathVertex>> vertexStore_;

  struct Parsed;
  void parseConfig(const std::string& specText);
  TopologyAwareScheduler* newTopologyAwareScheduler(const ChainCellList& ccl,
                                                   const std::map<int32_t, int32_t>& levelLeafCellNum,
                                                   bool crossPriorityPack);
  void initCellNums();
  void initPinnedCells(const std::map<std::string, std::map<std::string, Cell*>>& pinnedPcl);
  void initBadNodes();
  void assignIds();

  void setBadCell(Cell* c);
  void setHealthyCell(Cell* c);
  void addBadFreeCell(Cell* c);
  void removeBadFreeCell(Cell* c);
  void tryBindDoomedBadCell(const std::string& chain, int32_t l);
  void tryUnbindDoomedBadCell(const std::string& chain, int32_t l);

  void schedulePodFromExistingGroup(Group* g, const PodSchedulingSpec& s,
                                    const std::unordered_set<std::string>& suggestedNodes,
                                    bool preemptingPhase, int32_t podId, ScheduleResult& r);
  void schedulePodFromNewGroup(const PodSchedulingSpec& s,
                               const std::unordered_set<std::string>& suggestedNodes,
                               bool preemptingPhase, int32_t podId, ScheduleResult& r);
  void scheduleNewAffinityGroup(const PodSchedulingSpec& s,
                                const std::unordered_set<std::string>& suggestedNodes,
                                Placement& phys, Placement& virt, WaitReason& reason);
  void scheduleAffinityGroupForLeafCellType(SchedulingRequest& sr, const std::string& leafCellType,
                                            bool typeSpecified, Placement& phys, Placement& virt,
                                            WaitReason& reason);
  void scheduleAffinityGroupForAnyLeafCellType(SchedulingRequest& sr, Placement& phys,
                                               Placement& virt, WaitReason& reason);
  void validateSchedulingRequest(const SchedulingRequest& sr);
  void handleSchedulingRequest(SchedulingRequest& sr, Placement& phys, Placement& virt,
                               WaitReason& reason);
  void scheduleGuaranteedAffinityGroup(SchedulingRequest& sr, Placement& phys, Placement& virt,
                                       WaitReason& reason);
  std::map<std::string, Placement> tryLazyPreempt(const Placement& p,
                                                  const std::vector<int32_t>& leafCellNums,
                                                  const std::string& groupName);
  void scheduleOpportunisticAffinityGroup(SchedulingRequest& sr, Placement& phys,
                                          WaitReason& reason);
  void createAllocatedAffinityGroup(const PodSchedulingSpec& s, const PodBindInfo& info);
  void deleteAllocatedAffinityGroup(Group* g);
  void createPreemptingAffinityGroup(const PodSchedulingSpec& s, const Placement& phys,
                                     const Placement& virt, int32_t podId);
  void deletePreemptingAffinityGroup(Group* g);
  void allocatePreemptingAffinityGroup(Group* g);
  Placement lazyPreemptAffinityGroup(Group* victim, const std::string& preemptor);
  void lazyPreemptCell(Cell* c, const std::string& preemptor);
  void revertLazyPreempt(Group* g, const Placement& virtualPlacement);
  // lazyPreempt: 0 = nil, 1 = false, 2 = true
  void findAllocatedLeafCell(int32_t index, const std::vector<int32_t>& physicalLeafCellIndices,
                             const PodPlacementInfo& pp, const std::string& chain,
                             const std::string& node, bool lazyPreempted, const PodSchedulingSpec& s,
                             Group* group, Cell*& pLeafCell, Cell*& vLeafCell, int& lazyPreempt);
  bool allocateLeafCell(Cell* pLeafCell, Cell* vLeafCell, int32_t p, const std::string& vcn);
  void releaseLeafCell(Cell* pLeafCell, const std::string& vcn);
  bool allocatePreassignedCell(Cell* c, const std::string& vcn, bool doomedBad);
  void allocateBadCell(Cell* c);
  void releasePreassignedCell(Cell* c, const std::string& vcn, bool doomedBad);
  void releaseBadCell(Cell* c);
  int32_t removeCellFromFreeList(Cell* c);
  int32_t addCellToFreeList(Cell* c);

  // intra_vc_scheduler.go:92-117
  void intraVCSchedule(IntraVCScheduler* s, SchedulingRequest& sr, Placement& placement,
                       WaitReason& reason);
  // topology_aware_scheduler.go
  void tasSchedule(TopologyAwareScheduler* t, const std::map<int32_t, int32_t>& podLeafCellNumbers,
                   int32_t p, const std::unordered_set<std::string>& suggestedNodes,
                   bool ignoreSuggestedNodes, Placement& placement, WaitReason& reason);

  // cell_allocation.go
  bool buddyAlloc(BindingPathVertex* cell, ChainCellList& freeList, int32_t currentLevel,
                  const std::unordered_set<std::string>& suggestedNodes, bool ignoreSuggestedNodes,
                  std::map<std::string, Cell*>& bindings);
  bool safeRelaxedBuddyAlloc(BindingPathVertex* cell, ChainCellList& freeList,
                             std::map<int32_t, int32_t>& freeCellNum, int32_t currentLevel,
                             const std::unordered_set<std::string>& suggestedNodes,
                             bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings);
  bool mapVirtualPlacementToPhysical(std::vector<BindingPathVertex*>& preassignedCells,
                                     std::vector<std::vector<BindingPathVertex*>>& nonPreassignedCells,
                                     ChainCellList& freeList, std::map<int32_t, int32_t>& freeCellNum,
                                     const std::unordered_set<std::string>& suggestedNodes,
                                     bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings);
  bool getUsablePhysicalCells(const CellList& candidates, int32_t numNeeded,
                              const std::unordered_set<std::string>& suggestedNodes,
                              bool ignoreSuggestedNodes, CellList& usable);
  bool mapVirtualCellsToPhysical(const std::vector<BindingPathVertex*>& cells,
                                 const CellList& candidates,
                                 const std::unordered_set<std::string>& suggestedNodes,
                                 bool ignoreSuggestedNodes, std::map<std::string, Cell*>& bindings,
                                 bool returnPicked, CellList& pickedCells);
  // types.go:285-340
  void toBindingPaths(const Placement& p, const std::vector<int32_t>& leafCellNums,
                      std::map<std::string, Cell*>& bindings,
                      std::vector<BindingPathVertex*>& preassignedCells,
                      std::vector<std::vector<BindingPathVertex*>>& nonPreassignedCells);
  BindingPathVertex* newVertex(Cell* c);

  // utils.go
  void generatePodScheduleResult(ScheduleResult& r, int32_t currentLeafCellNum,
                                 int32_t currentPodIndex, Group* group, const std::string& groupName);
  Cell* findPhysicalLeafCell(const std::string& chain, const std::string& node, int32_t leafCellIndex);
  Cell* findPhysicalLeafCellInChain(const std::string& chain, const std::string& node,
                                    int32_t leafCellIndex);
};

// free functions of the reference that tests may want to poke
bool inFreeCellList(Cell* c);                     // utils.go:381-391
void setCellState(Cell* c, int32_t s);            // utils.go:397-405
void setCellPriority(Cell* c, int32_t p);         // cell_allocation.go:425-441
Cell* findLCA(Cell* lower, Cell* higher);         // topology_aware_scheduler.go:444-462

}  // namespace hived_oracle
