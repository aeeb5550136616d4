// Native network runtime: a descriptor array (ConvPlan) compiled into a CUDA graph whose independent branches overlap.
//
// The reference runs its encoder / UNet++ decoder (modules/networks.py:65-117) as ~170 dependent cuDNN launches in Python
// program order.  Most of those layers are NOT dependent on each other: in decoder column j every right_conv / diag_conv
// block reads only column j-1, and a BasicBlock's 1x1 skip projection is independent of its conv1 (modules/layers.py:77-94).
// The tensor-core conv kernel is persistent (one CTA per SM, static tile striding), so a layer whose tile count is not a
// multiple of 148 leaves most SMs idle during its last round (120x160 maps: 150 tiles = 2 rounds for 1.01 rounds of
// work).  Running independent layers concurrently lets the next layer's CTAs start on those SMs immediately.
//
//   analyse   host only: byte ranges every op reads / writes  ->  RAW / WAW / WAR edges  ->  transitive reduction,
//             lane (capture stream) assignment that never adds a false dependency while lanes are available
//   capture   replay the existing per-op launchers onto the lanes under stream capture (events become graph edges)
//   launch    one cudaGraphLaunch per network per frame
#include <vector>

#include "common.cuh"

namespace dtb200 {

uint64_t conv_tc_workspace_bytes(const dtb200_conv_params& p, int in_c_total);  // conv_tc.cu
int conv_tc_init();                                                             // conv_tc.cu
int conv_tch_init();                                                            // conv_tch.cu

namespace {

struct Range {
  uintptr_t lo, hi;
};
inline bool overlap(const Range& a, const Range& b) { return a.lo < b.hi && b.lo < a.hi; }

struct OpAccess {
  Range reads[DTB200_CONV_MAX_SRC + 1];
  int num_reads = 0;
  Range writes[2];
  int num_writes = 0;
};

inline Range range_of(const void* p, uint64_t bytes) {
  uintptr_t lo = reinterpret_cast<uintptr_t>(p);
  return Range{lo, lo + bytes};
}

OpAccess access_of(const dtb200_conv_params& p) {
  OpAccess a;
  for (int s = 0; s < p.num_src && s < DTB200_CONV_MAX_SRC; ++s) {
    if (!p.src[s]) continue;
    const bool half = p.src_resample[s] != DTB200_RESAMPLE_NONE;
    const uint64_t h = half ? p.in_h / 2 : p.in_h, w = half ? p.in_w / 2 : p.in_w;
    a.reads[a.num_reads++] = range_of(p.src[s], (uint64_t)p.batch * h * w * p.src_c[s] * sizeof(float));
  }
  const uint64_t out_bytes = (uint64_t)p.batch * p.out_h * p.out_w * p.out_c * sizeof(float);
  if (p.residual && p.ksize != 0) a.reads[a.num_reads++] = range_of(p.residual, out_bytes);
  if (p.dst) a.writes[a.num_writes++] = range_of(p.dst, out_bytes);
  if (p.workspace && p.ksize != 0) {
    const uint64_t ws = dtb200_conv_workspace_bytes(&p);   // split-K scratch of the tensor-core modes
    if (ws) a.writes[a.num_writes++] = range_of(p.workspace, ws);
  }
  return a;
}

bool conflicts(const OpAccess& earlier, const OpAccess& later) {
  for (int i = 0; i < earlier.num_writes; ++i) {
    for (int j = 0; j < later.num_reads; ++j)
      if (overlap(earlier.writes[i], later.reads[j])) return true;  // RAW
    for (int j = 0; j < later.num_writes; ++j)
      if (overlap(earlier.writes[i], later.writes[j])) return true;  // WAW
  }
  for (int i = 0; i < earlier.num_reads; ++i)
    for (int j = 0; j < later.num_writes; ++j)
      if (overlap(earlier.reads[i], later.writes[j])) return true;  // WAR
  return false;
}

struct Bitset {
  std::vector<uint64_t> w;
  explicit Bitset(int n = 0) : w((n + 63) / 64, 0) {}
  void set(int i) { w[i >> 6] |= 1ull << (i & 63); }
  bool test(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
  void merge(const Bitset& o) {
    for (size_t i = 0; i < w.size(); ++i) w[i] |= o.w[i];
  }
};

struct Schedule {
  int count = 0, lanes_used = 0, depth = 0, edges = 0;
  std::vector<int> lane_of, level_of;
  std::vector<std::vector<int>> deps;   // direct dependencies after transitive reduction
  std::vector<std::vector<int>> waits;  // dependencies that need an event wait (not implied by lane order)
};

// Ops arrive in a valid serial order (the plan builder's program order); edges only point backwards.
Schedule analyse(const dtb200_conv_params* ops, int count, int max_lanes) {
  Schedule sc;
  sc.count = count;
  sc.lane_of.assign(count, 0);
  sc.level_of.assign(count, 0);
  sc.deps.resize(count);
  sc.waits.resize(count);
  if (max_lanes < 1) max_lanes = 1;
  std::vector<OpAccess> acc(count);
  for (int i = 0; i < count; ++i) acc[i] = access_of(ops[i]);
  std::vector<Bitset> anc(count, Bitset(count));  // strict ancestors
  std::vector<int> last_on_lane;                  // last op issued on each lane
  for (int i = 0; i < count; ++i) {
    std::vector<int> all;
    for (int j = 0; j < i; ++j)
      if (conflicts(acc[j], acc[i])) all.push_back(j);
    // transitive reduction: drop j when another dependency already has j as an ancestor
    for (int j : all) {
      bool implied = false;
      for (int k : all)
        if (k != j && anc[k].test(j)) {
          implied = true;
          break;
        }
      if (!implied) sc.deps[i].push_back(j);
    }
    for (int j : all) {
      anc[i].set(j);
      anc[i].merge(anc[j]);
    }
    int level = 0;
    for (int j : sc.deps[i]) level = level > sc.level_of[j] + 1 ? level : sc.level_of[j] + 1;
    sc.level_of[i] = level;
    if (level + 1 > sc.depth) sc.depth = level + 1;
    sc.edges += (int)sc.deps[i].size();
    // lane: continue the chain of a direct dependency; else sit behind an ancestor (no false edge); else a new lane;
    // else the lane whose last op is shallowest (a false edge, only when every lane is busy with unrelated work)
    int lane = -1;
    for (int l = 0; l < (int)last_on_lane.size() && lane < 0; ++l)
      for (int j : sc.deps[i])
        if (last_on_lane[l] == j) {
          lane = l;
          break;
        }
    if (lane < 0)
      for (int l = 0; l < (int)last_on_lane.size(); ++l)
        if (anc[i].test(last_on_lane[l])) {
          lane = l;
          break;
        }
    if (lane < 0 && (int)last_on_lane.size() < max_lanes) {
      lane = (int)last_on_lane.size();
      last_on_lane.push_back(-1);
    }
    if (lane < 0) {
      lane = 0;
      for (int l = 1; l < (int)last_on_lane.size(); ++l)
        if (sc.level_of[last_on_lane[l]] < sc.level_of[last_on_lane[lane]]) lane = l;
    }
    const int prev = last_on_lane[lane];
    for (int j : sc.deps[i]) {
      const bool implied_by_lane = prev >= 0 && (j == prev || anc[prev].test(j));
      if (!implied_by_lane) sc.waits[i].push_back(j);
    }
    sc.lane_of[i] = lane;
    last_on_lane[lane] = i;
  }
  sc.lanes_used = (int)last_on_lane.size();
  return sc;
}

}  // namespace
}  // namespace dtb200

using namespace dtb200;

struct dtb200_conv_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int nodes = 0, edges = 0, lanes = 0, depth = 0, ops = 0;
  uint64_t kernels = 0;  // kernel launches captured (added to the launch counter on every replay)
};

extern "C" int dtb200_conv_graph_analyze(const dtb200_conv_params* ops, int32_t count, int32_t max_lanes, int32_t* lane_of,
                                         int32_t* level_of, int32_t* dep_offsets, int32_t* deps, int32_t deps_capacity) {
  if (!ops || count < 0 || max_lanes < 1) return fail(DTB200_ERR_INVALID, "conv graph analyze: bad arguments%s");
  Schedule sc = analyse(ops, count, max_lanes);
  int n = 0;
  for (int i = 0; i < count; ++i) {
    if (lane_of) lane_of[i] = sc.lane_of[i];
    if (level_of) level_of[i] = sc.level_of[i];
    if (dep_offsets) dep_offsets[i] = n;
    for (int j : sc.deps[i]) {
      if (deps && n < deps_capacity) deps[n] = j;
      ++n;
    }
  }
  if (dep_offsets) dep_offsets[count] = n;
  if (deps && n > deps_capacity) return fail(DTB200_ERR_INVALID, "conv graph analyze: deps_capacity too small (%s%lld needed)", "", n);
  return DTB200_OK;
}

extern "C" void dtb200_conv_graph_destroy(dtb200_conv_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

extern "C" int dtb200_conv_graph_create(const dtb200_conv_params* ops, int32_t count, int32_t max_lanes,
                                        dtb200_conv_graph** out) {
  if (!ops || count < 1 || !out || max_lanes < 1) return fail(DTB200_ERR_INVALID, "conv graph: bad arguments%s");
  *out = nullptr;
  Schedule sc = analyse(ops, count, max_lanes);
  conv_tc_init();  // resolve driver entry points / device attributes before capture starts
  conv_tch_init();

  std::vector<cudaStream_t> lanes(sc.lanes_used, nullptr);
  std::vector<cudaEvent_t> done(count, nullptr), joined(sc.lanes_used, nullptr);
  cudaEvent_t start = nullptr;
  cudaGraph_t graph = nullptr;
  int rc = DTB200_OK;
  bool capturing = false;
  auto cuda_ok = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    rc = fail(DTB200_ERR_CUDA, "conv graph: %s", what);
    snprintf(g_error, sizeof(g_error), "conv graph: %s: %s", what, cudaGetErrorString(e));
    return false;
  };
  do {
    bool ok = true;
    for (auto& s : lanes) ok = ok && cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "stream create");
    for (auto& e : done) ok = ok && cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create");
    for (auto& e : joined) ok = ok && cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event create");
    ok = ok && cuda_ok(cudaEventCreateWithFlags(&start, cudaEventDisableTiming), "event create");
    if (!ok) break;
    if (!cuda_ok(cudaStreamBeginCapture(lanes[0], cudaStreamCaptureModeRelaxed), "begin capture")) break;
    capturing = true;
    if (!cuda_ok(cudaEventRecord(start, lanes[0]), "record start")) break;
    std::vector<char> lane_joined(sc.lanes_used, 0);
    lane_joined[0] = 1;
    const uint64_t launches_before = g_launches.load();
    for (int i = 0; i < count && rc == DTB200_OK; ++i) {
      cudaStream_t s = lanes[sc.lane_of[i]];
      if (!lane_joined[sc.lane_of[i]]) {  // a lane enters the capture by waiting on an event of a capturing stream
        if (!cuda_ok(cudaStreamWaitEvent(s, start, 0), "fork lane")) break;
        lane_joined[sc.lane_of[i]] = 1;
      }
      for (int j : sc.waits[i])
        if (!cuda_ok(cudaStreamWaitEvent(s, done[j], 0), "wait dependency")) break;
      if (rc != DTB200_OK) break;
      rc = dtb200_conv2d(&ops[i], s);
      if (rc != DTB200_OK) break;
      if (!cuda_ok(cudaEventRecord(done[i], s), "record op")) break;
    }
    if (rc != DTB200_OK) break;
    const uint64_t kernels = g_launches.load() - launches_before;
    for (int l = 1; l < sc.lanes_used && rc == DTB200_OK; ++l) {
      if (!lane_joined[l]) continue;
      if (!cuda_ok(cudaEventRecord(joined[l], lanes[l]), "record join")) break;
      if (!cuda_ok(cudaStreamWaitEvent(lanes[0], joined[l], 0), "join lane")) break;
    }
    if (rc != DTB200_OK) break;
    capturing = false;
    if (!cuda_ok(cudaStreamEndCapture(lanes[0], &graph), "end capture")) break;
    dtb200_conv_graph* g = new dtb200_conv_graph();
    g->graph = graph;
    graph = nullptr;
    if (!cuda_ok(cudaGraphInstantiate(&g->exec, g->graph, 0), "instantiate")) {
      dtb200_conv_graph_destroy(g);
      break;
    }
    size_t nodes = 0, edges = 0;
    cudaGraphGetNodes(g->graph, nullptr, &nodes);
    cudaGraphGetEdges(g->graph, nullptr, nullptr, &edges);
    g->nodes = (int)nodes;
    g->edges = (int)edges;
    g->lanes = sc.lanes_used;
    g->depth = sc.depth;
    g->ops = count;
    g->kernels = kernels;
    *out = g;
  } while (false);
  if (capturing) {  // leave capture mode so the streams can be destroyed; the partial graph is discarded
    cudaGraph_t partial = nullptr;
    cudaStreamEndCapture(lanes[0], &partial);
    if (partial) cudaGraphDestroy(partial);
    cudaGetLastError();
  }
  if (graph) cudaGraphDestroy(graph);
  for (auto e : done)
    if (e) cudaEventDestroy(e);
  for (auto e : joined)
    if (e) cudaEventDestroy(e);
  if (start) cudaEventDestroy(start);
  for (auto s : lanes)
    if (s) cudaStreamDestroy(s);
  return rc;
}

extern "C" int dtb200_conv_graph_launch(dtb200_conv_graph* g, dtb200_stream_t stream) {
  if (!g || !g->exec) return fail(DTB200_ERR_INVALID, "conv graph launch: null graph%s");
  cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    snprintf(g_error, sizeof(g_error), "cudaGraphLaunch: %s", cudaGetErrorString(e));
    return DTB200_ERR_CUDA;
  }
  g_launches.fetch_add(g->kernels, std::memory_order_relaxed);
  return DTB200_OK;
}

extern "C" int dtb200_conv_graph_info(const dtb200_conv_graph* g, int32_t* ops, int32_t* kernel_nodes, int32_t* edges,
                                      int32_t* lanes, int32_t* depth) {
  if (!g) return fail(DTB200_ERR_INVALID, "conv graph info: null graph%s");
  if (ops) *ops = g->ops;
  if (kernel_nodes) *kernel_nodes = (int32_t)g->kernels;
  if (edges) *edges = g->edges;
  if (lanes) *lanes = g->lanes;
  if (depth) *depth = g->depth;
  return DTB200_OK;
}
