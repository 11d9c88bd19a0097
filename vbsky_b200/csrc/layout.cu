// Tree layout: validation of the reference's TreeData arrays and construction of the pruning
// kernel's evaluation schedules (host), plus upload.  See include/vbsky_b200.h and DESIGN.md.
//
// The reference walks parents in index order n..2n-2 (prune.py:49) and children in reversed
// postorder (prune.py:95), materialising every partial vector.  Per-node results do not depend
// on the visiting order, so the schedules here are chosen to keep the number of simultaneously
// live vectors minimal (a Strahler / Sethi-Ullman ordering, <= log2(n)+1 for any binary tree),
// which lets the kernel hold them in shared memory.
#include <algorithm>
#include <cstring>

#include "common.h"

namespace vbsky {

static thread_local std::string g_err;

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

namespace {

struct SlotPool {
  std::vector<char> used;
  int high = 0;
  int alloc() {
    for (size_t i = 0; i < used.size(); ++i)
      if (!used[i]) {
        used[i] = 1;
        high = std::max(high, (int)i + 1);
        return (int)i;
      }
    used.push_back(1);
    high = std::max(high, (int)used.size());
    return (int)used.size() - 1;
  }
  void release(int s) { used[s] = 0; }
};

// Build both schedules of one tree.  pc = parent_child [N][2].
int build_tree(int n, const int32_t* pc, Op* up, Op* down, int32_t* down_nodes, int32_t* msg_rows,
               int32_t* msg_start, int32_t* dchunk_start, int32_t* n_dchunks, int32_t* n_msgs, Op* chunk_meta, int* depth_up,
               int* depth_down) {
  const int N = 2 * n - 1, root = N - 1;
  auto internal = [n](int v) { return v >= n; };
  // A "cherry" is a non-root internal node whose two children are tips.  Its message is cheap to recompute from the two
  // tip tables (two gathers, a product and one 4x4 product per site), so the up pass does not stream it to scratch and
  // the down pass rebuilds it at the parent's step from the cherry's own down-record, which the schedule places right
  // after the parent's (same ring stage): about a third of all messages of a random tree never touch HBM.
  auto cherry = [&](int v) { return v >= n && v != root && pc[2 * v] < n && pc[2 * v + 1] < n; };

  // ------------------------------------------------------------------------------------ up pass
  // need[v] = peak number of messages parked in shared memory while evaluating subtree v.  The message
  // produced by a step is carried in registers into the next step, so only a first-evaluated internal
  // child waits in a slot while its sibling's subtree is evaluated.
  std::vector<int> need(N, 0), first(N, 0);  // first[v]: which child (0/1) is evaluated first
  for (int v = n; v < N; ++v) {
    const int c[2] = {pc[2 * v], pc[2 * v + 1]};
    int best = 1 << 30, bestf = 0;
    for (int f = 0; f < 2; ++f) {
      const int a = c[f], b = c[1 - f];
      // a's result must be parked only if something is evaluated between a and v, i.e. b is internal
      const int park = (internal(a) && internal(b)) ? 1 : 0;
      const int peak = std::max(need[a], park + need[b]);
      if (peak < best) best = peak, bestf = f;
    }
    need[v] = best;
    first[v] = bestf;
  }
  std::vector<int> order;  // nodes in evaluation order
  order.reserve(n - 1);
  {
    std::vector<std::pair<int, int>> st;
    st.push_back({root, 0});
    while (!st.empty()) {
      const int v = st.back().first;
      const int a = pc[2 * v + first[v]], b = pc[2 * v + 1 - first[v]];
      if (st.back().second == 0) {
        st.back().second = 1;
        if (internal(a)) {
          st.push_back({a, 0});
          continue;
        }
      }
      if (st.back().second == 1) {
        st.back().second = 2;
        if (internal(b)) {
          st.push_back({b, 0});
          continue;
        }
      }
      order.push_back(v);
      st.pop_back();
    }
  }
  if ((int)order.size() != n - 1) return set_error(VBSKY_EINVAL, "internal error: up schedule has %zu steps, expected %d", order.size(), n - 1);
  std::vector<int> up_row(N, -1), slot_of(N, -1);
  for (int s = 0; s < n - 1; ++s) up_row[order[s]] = s;
  std::vector<int> parent(N, -1);
  for (int v = n; v < N; ++v) parent[pc[2 * v]] = v, parent[pc[2 * v + 1]] = v;
  {
    SlotPool pool;
    for (int s = 0; s < n - 1; ++s) {
      const int v = order[s];
      const int a = pc[2 * v + first[v]], b = pc[2 * v + 1 - first[v]];
      int so = 0xff, flags = 0;
      // source of each child's message: 0 tip, 1 previous step's result (registers), 2 parked in a slot
      int ch[2] = {a, b}, src[2] = {0, 0}, slot[2] = {0xff, 0xff};
      for (int k = 0; k < 2; ++k) {
        const int c = ch[k];
        if (!internal(c)) {
          src[k] = 0;
        } else if (up_row[c] == s - 1) {
          src[k] = 1;
        } else {
          src[k] = 2;
          slot[k] = slot_of[c];
          pool.release(slot_of[c]);
        }
      }
      if (src[1] == 1) std::swap(ch[0], ch[1]), std::swap(src[0], src[1]), std::swap(slot[0], slot[1]);  // canonical: only a is ever carried
      for (int k = 0; k < 2; ++k) {
        if (src[k] == 0) flags |= 1 << k;
        if (src[k] == 1) flags |= 8 << k;
      }
      const int ca = ch[0], cb = ch[1], sa = slot[0], sb = slot[1];
      if (cherry(v)) flags |= 0x40;  // message is not written to scratch
      if (v == root) {
        flags |= 4;
      } else if (up_row[parent[v]] != s + 1) {
        so = pool.alloc();  // parked until the sibling's subtree is done
        slot_of[v] = so;
        flags |= 0x20;
        if (so >= 0xff) return set_error(VBSKY_EINVAL, "tree needs more than 254 live vectors");
      }
      up[s] = Op{ca | (cb << 16), v, sa | (sb << 8) | (so << 16), flags};
    }
    *depth_up = std::max(pool.high, 1);
  }

  // ---------------------------------------------------------------------------------- down pass
  // needd[v] = peak number of pre vectors parked while expanding subtree v.  The pre vector of the child
  // that is expanded next is carried in registers; the other internal child's waits in a slot.
  std::vector<int> needd(N, 0), firstd(N, 0);
  for (int v = n; v < N; ++v) {
    const int a = pc[2 * v], b = pc[2 * v + 1];
    const int ia = internal(a), ib = internal(b);
    if (!ia && !ib)
      needd[v] = 0;
    else if (ia && !ib)
      needd[v] = needd[a], firstd[v] = 0;
    else if (!ia && ib)
      needd[v] = needd[b], firstd[v] = 1;
    else {
      // descend first into the child with the smaller need while the other's pre vector waits; a cherry always goes
      // first (its need is 0, so this is never worse) so that its step directly follows the parent's
      const int f = cherry(a) ? 0 : (cherry(b) ? 1 : (needd[a] <= needd[b] ? 0 : 1));
      const int cf = pc[2 * v + f], cs = pc[2 * v + 1 - f];
      needd[v] = std::max(needd[cf] + 1, needd[cs]);
      firstd[v] = f;
    }
  }
  std::vector<int> dorder;
  dorder.reserve(n - 1);
  {
    std::vector<int> st;
    st.push_back(root);
    while (!st.empty()) {
      const int v = st.back();
      st.pop_back();
      dorder.push_back(v);
      const int a = pc[2 * v + firstd[v]], b = pc[2 * v + 1 - firstd[v]];
      if (internal(b)) st.push_back(b);
      if (internal(a)) st.push_back(a);  // a is expanded next
    }
  }
  if ((int)dorder.size() != n - 1) return set_error(VBSKY_EINVAL, "internal error: down schedule has %zu steps", dorder.size());
  {
    SlotPool pool;
    std::fill(slot_of.begin(), slot_of.end(), -1);
    int nmsg = 0, nchunk = 0, fill = 0;
    // Ring stages hold up to kChunkSteps consecutive steps; a step and the cherry children it rebuilds (the next one or
    // two steps) are never split across stages, so chunks have variable length.
    auto group_size = [&](int v) { return 1 + (cherry(pc[2 * v]) ? 1 : 0) + (cherry(pc[2 * v + 1]) ? 1 : 0); };
    for (int s = 0; s < n - 1; ++s) {
      const int v = dorder[s];
      const bool member = cherry(v);  // belongs to its parent's group (already accounted for)
      if (!member) {
        const int g = group_size(v);
        if (s == 0 || fill + g > kChunkSteps) {
          dchunk_start[nchunk] = s, msg_start[nchunk] = nmsg, ++nchunk, fill = 0;
        }
        fill += g;
      }
      int a = pc[2 * v + firstd[v]], b = pc[2 * v + 1 - firstd[v]];
      if (!internal(a) && internal(b)) std::swap(a, b);  // canonical: an internal a is expanded next (carried), an internal b is parked
      const int next = s + 1 < n - 1 ? dorder[s + 1] : -1;
      int sp = 0xff, sa = 0xff, sb = 0xff, flags = 0;
      if (v == root) {
        flags |= 8;  // pre[root] = pi: the consumer seeds the carry registers with it before the walk
      } else if (s > 0 && parent[v] == dorder[s - 1]) {
        flags |= 8;  // carried in registers from the previous step
      } else {
        sp = slot_of[v];
        pool.release(sp);
      }
      if (internal(a) && a != next) return set_error(VBSKY_EINVAL, "internal error: down schedule order");
      if (!internal(a)) flags |= 1;
      else if (cherry(a)) flags |= 0x40;  // rebuilt from the next record
      else msg_rows[nmsg++] = up_row[a];
      if (!internal(b)) {
        flags |= 2;
      } else {
        if (cherry(b)) {
          // only reachable when a is a cherry too (a cherry always goes first): b's step is the one after a's
          if (!cherry(a) || s + 2 >= n - 1 || dorder[s + 2] != b) return set_error(VBSKY_EINVAL, "internal error: cherry order");
          flags |= 0x80;  // rebuilt from the record after next
        } else {
          msg_rows[nmsg++] = up_row[b];
        }
        sb = pool.alloc();  // parked until the walk returns to it
        slot_of[b] = sb;
        if (sb >= 0xff) return set_error(VBSKY_EINVAL, "tree needs more than 254 live vectors");
      }
      down[s] = Op{a | (b << 16), v, sp | (sa << 8) | (sb << 16), flags};
      down_nodes[2 * s] = a;
      down_nodes[2 * s + 1] = b;
    }
    dchunk_start[nchunk] = n - 1, msg_start[nchunk] = nmsg;
    *n_dchunks = nchunk, *n_msgs = nmsg;
    // one 16-byte record per chunk for the producer lane: {first step, steps, first message, end message}
    for (int c = 0; c < nchunk; ++c)
      chunk_meta[c] = Op{dchunk_start[c], dchunk_start[c + 1] - dchunk_start[c], msg_start[c], msg_start[c + 1]};
    *depth_down = std::max(pool.high, 1);
  }
  return VBSKY_OK;
}

template <class T>
int upload(T** dptr, const std::vector<T>& h) {
  VB_CUDA(cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) VB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return VBSKY_OK;
}

}  // namespace
}  // namespace vbsky

using namespace vbsky;

extern "C" const char* vbsky_last_error(void) { return g_err.c_str(); }
extern "C" int vbsky_version(void) { return 100; }
extern "C" int vbsky_device_count(void) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) return set_error(VBSKY_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return c;
}

extern "C" void vbsky_layout_destroy(vbsky_layout* lay) {
  if (!lay) return;
  cudaFree(lay->d_up);
  cudaFree(lay->d_down);
  cudaFree(lay->d_down_nodes);
  cudaFree(lay->d_msg_rows);
  cudaFree(lay->d_msg_start);
  cudaFree(lay->d_dchunk_start);
  cudaFree(lay->d_n_dchunks);
  cudaFree(lay->d_n_msgs);
  cudaFree(lay->d_chunk_meta);
  cudaFree(lay->d_child_parent);
  cudaFree(lay->d_parent_child);
  cudaFree(lay->d_lst);
  cudaFree(lay->d_sample_times);
  cudaFree(lay->d_max_sample_time);
  cudaFree(lay->d_level_nodes);
  cudaFree(lay->d_level_start);
  delete lay;
}

extern "C" int vbsky_layout_create(int32_t T, int32_t n, const int32_t* parent_child,
                                   const int32_t* child_parent, const int32_t* postorder,
                                   const double* sample_times, vbsky_layout** out) {
  VB_CHECK_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  VB_CHECK_ARG(T >= 1, "T must be >= 1 (got %d)", T);
  VB_CHECK_ARG(n >= 2 && n <= 32767, "n must be in [2, 32767] (got %d)", n);
  VB_CHECK_ARG(parent_child && child_parent && postorder && sample_times, "NULL input array");
  const int N = 2 * n - 1, root = N - 1;

  auto* lay = new vbsky_layout();
  lay->T = T, lay->n = n, lay->N = N;
  lay->h_up.resize((size_t)T * (n - 1));
  lay->h_down.resize((size_t)T * (n - 1));
  lay->h_down_nodes.resize((size_t)T * (2 * n - 2));
  const int chunk_cap = n - 1;  // variable-length down chunks: at most one per step
  lay->chunk_cap = chunk_cap;
  lay->h_msg_rows.assign((size_t)T * std::max(n - 2, 1), 0);
  lay->h_msg_start.assign((size_t)T * (chunk_cap + 1), 0);
  lay->h_dchunk_start.assign((size_t)T * (chunk_cap + 1), 0);
  lay->h_n_dchunks.assign(T, 0);
  lay->h_n_msgs.assign(T, 0);
  lay->h_chunk_meta.assign((size_t)T * std::max(chunk_cap, 1), Op{0, 0, 0, 0});
  lay->h_child_parent.assign(child_parent, child_parent + (size_t)T * N);
  lay->h_parent_child.assign(parent_child, parent_child + (size_t)T * N * 2);
  lay->h_siblings.assign((size_t)T * N, -1);
  lay->h_lst.assign((size_t)T * N, 0.0);
  lay->h_sample_times.assign(sample_times, sample_times + (size_t)T * n);
  lay->h_max_sample_time.resize(T);
  lay->h_level_nodes.resize((size_t)T * (n - 1));
  std::vector<std::vector<int32_t>> level_starts(T);

  int rc = VBSKY_OK;
  for (int t = 0; t < T && rc == VBSKY_OK; ++t) {
    const int32_t* pc = parent_child + (size_t)t * N * 2;
    const int32_t* cp = child_parent + (size_t)t * N;
    const int32_t* po = postorder + (size_t)t * N;
    // validation (the reference relies on these silently: prune.py:49, bdsky.py:307)
    std::vector<char> seen(N, 0);
    for (int i = 0; i < N && rc == VBSKY_OK; ++i) {
      if (po[i] < 0 || po[i] >= N || seen[po[i]]) rc = set_error(VBSKY_EINVAL, "tree %d: postorder is not a permutation of 0..%d", t, N - 1);
      else seen[po[i]] = 1;
    }
    if (rc == VBSKY_OK && po[N - 1] != root) rc = set_error(VBSKY_EINVAL, "tree %d: postorder must end at the root %d", t, root);
    if (rc == VBSKY_OK && cp[root] != -1) rc = set_error(VBSKY_EINVAL, "tree %d: child_parent[root] must be -1", t);
    for (int v = 0; v < N && rc == VBSKY_OK; ++v) {
      const int a = pc[2 * v], b = pc[2 * v + 1];
      if (v < n) {
        if (a != -1 || b != -1) rc = set_error(VBSKY_EINVAL, "tree %d: leaf %d has children", t, v);
      } else {
        if (a < 0 || b < 0 || a >= v || b >= v || a == b) rc = set_error(VBSKY_EINVAL, "tree %d: node %d needs two distinct children with smaller index (got %d, %d)", t, v, a, b);
        else if (cp[a] != v || cp[b] != v) rc = set_error(VBSKY_EINVAL, "tree %d: parent_child and child_parent disagree at node %d", t, v);
      }
      if (rc == VBSKY_OK && v != root && (cp[v] < n || cp[v] >= N)) rc = set_error(VBSKY_EINVAL, "tree %d: node %d has invalid parent %d", t, v, cp[v]);
    }
    if (rc != VBSKY_OK) break;

    int du = 0, dd = 0;
    rc = build_tree(n, pc, &lay->h_up[(size_t)t * (n - 1)], &lay->h_down[(size_t)t * (n - 1)],
                    &lay->h_down_nodes[(size_t)t * (2 * n - 2)], &lay->h_msg_rows[(size_t)t * std::max(n - 2, 1)],
                    &lay->h_msg_start[(size_t)t * (chunk_cap + 1)], &lay->h_dchunk_start[(size_t)t * (chunk_cap + 1)],
                    &lay->h_n_dchunks[t], &lay->h_n_msgs[t], &lay->h_chunk_meta[(size_t)t * std::max(chunk_cap, 1)], &du, &dd);
    if (rc != VBSKY_OK) break;
    lay->depth_up = std::max(lay->depth_up, du);
    lay->depth_down = std::max(lay->depth_down, dd);

    // siblings (tree_data.py:68-78) and lower sampling times (tree_data.py:80-117)
    int32_t* sib = &lay->h_siblings[(size_t)t * N];
    double* lst = &lay->h_lst[(size_t)t * N];
    const double* stimes = sample_times + (size_t)t * n;
    double mx = stimes[0];
    for (int i = 0; i < n; ++i) lst[i] = stimes[i], mx = std::max(mx, stimes[i]);
    lay->h_max_sample_time[t] = mx;
    for (int v = n; v < N; ++v) {
      const int a = pc[2 * v], b = pc[2 * v + 1];
      sib[a] = b, sib[b] = a;
      lst[v] = std::max(lst[a], lst[b]);
    }
    // levels of the internal nodes (root = level 0), parents before children
    std::vector<int> depth(N, 0);
    int maxd = 0;
    for (int v = root - 1; v >= n; --v) depth[v] = depth[cp[v]] + 1, maxd = std::max(maxd, depth[v]);
    std::vector<int32_t>& ls = level_starts[t];
    ls.assign(maxd + 2, 0);
    for (int v = n; v < N; ++v) ls[depth[v] + 1]++;
    for (int l = 0; l <= maxd; ++l) ls[l + 1] += ls[l];
    std::vector<int> fill(ls.begin(), ls.end() - 1);
    int32_t* ln = &lay->h_level_nodes[(size_t)t * (n - 1)];
    for (int v = n; v < N; ++v) ln[fill[depth[v]]++] = v;
    lay->max_levels = std::max(lay->max_levels, maxd + 1);
  }
  if (rc != VBSKY_OK) {
    delete lay;
    return rc;
  }
  lay->h_level_start.assign((size_t)T * (lay->max_levels + 1), 0);
  for (int t = 0; t < T; ++t) {
    int32_t* dst = &lay->h_level_start[(size_t)t * (lay->max_levels + 1)];
    const auto& ls = level_starts[t];
    for (int l = 0; l <= lay->max_levels; ++l) dst[l] = l < (int)ls.size() ? ls[l] : ls.back();
  }

  // Without a visible CUDA device the layout stays host-only (schedules can still be inspected
  // through vbsky_layout_get); every compute entry point then refuses it -- there is no CPU path.
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    lay->device = -1;
    *out = lay;
    return VBSKY_OK;
  }
  cudaError_t e = cudaGetDevice(&lay->device);
  if (e != cudaSuccess) {
    delete lay;
    return set_error(VBSKY_ECUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
  }
#define VB_UP(dst, src)                                 \
  if ((rc = upload(&lay->dst, lay->src)) != VBSKY_OK) { \
    vbsky_layout_destroy(lay);                          \
    return rc;                                          \
  }
  VB_UP(d_up, h_up)
  VB_UP(d_down, h_down)
  VB_UP(d_down_nodes, h_down_nodes)
  VB_UP(d_msg_rows, h_msg_rows)
  VB_UP(d_msg_start, h_msg_start)
  VB_UP(d_dchunk_start, h_dchunk_start)
  VB_UP(d_n_dchunks, h_n_dchunks)
  VB_UP(d_n_msgs, h_n_msgs)
  VB_UP(d_chunk_meta, h_chunk_meta)
  VB_UP(d_child_parent, h_child_parent)
  VB_UP(d_parent_child, h_parent_child)
  VB_UP(d_lst, h_lst)
  VB_UP(d_sample_times, h_sample_times)
  VB_UP(d_max_sample_time, h_max_sample_time)
  VB_UP(d_level_nodes, h_level_nodes)
  VB_UP(d_level_start, h_level_start)
#undef VB_UP
  *out = lay;
  return VBSKY_OK;
}

extern "C" int vbsky_layout_info(const vbsky_layout* lay, int32_t info[8]) {
  VB_CHECK_ARG(lay && info, "NULL argument");
  std::memset(info, 0, 8 * sizeof(int32_t));
  info[0] = lay->T, info[1] = lay->n, info[2] = lay->depth_up, info[3] = lay->depth_down, info[4] = lay->max_levels;
  info[5] = lay->chunk_cap;
  long long tot = 0;
  for (int v : lay->h_n_msgs) tot += v;
  info[6] = (int32_t)(tot / std::max(lay->T, 1));  // mean number of messages per tree that travel through scratch
  return VBSKY_OK;
}

extern "C" int vbsky_layout_get_chunks(const vbsky_layout* lay, int32_t t, int32_t* n_chunks, int32_t* chunk_start,
                                       int32_t* msg_start, int32_t* n_msgs, int32_t* msg_rows) {
  VB_CHECK_ARG(lay, "NULL layout");
  VB_CHECK_ARG(t >= 0 && t < lay->T, "tree index %d out of range", t);
  const int cap = lay->chunk_cap;
  if (n_chunks) *n_chunks = lay->h_n_dchunks[t];
  if (n_msgs) *n_msgs = lay->h_n_msgs[t];
  if (chunk_start) std::memcpy(chunk_start, &lay->h_dchunk_start[(size_t)t * (cap + 1)], sizeof(int32_t) * (cap + 1));
  if (msg_start) std::memcpy(msg_start, &lay->h_msg_start[(size_t)t * (cap + 1)], sizeof(int32_t) * (cap + 1));
  if (msg_rows) std::memcpy(msg_rows, &lay->h_msg_rows[(size_t)t * std::max(lay->n - 2, 1)], sizeof(int32_t) * std::max(lay->n - 2, 1));
  return VBSKY_OK;
}

extern "C" int vbsky_layout_get(const vbsky_layout* lay, int32_t t, int32_t* up, int32_t* down,
                                int32_t* down_nodes, double* lst, int32_t* siblings) {
  VB_CHECK_ARG(lay, "NULL layout");
  VB_CHECK_ARG(t >= 0 && t < lay->T, "tree index %d out of range", t);
  const int n = lay->n, N = lay->N;
  if (up) std::memcpy(up, &lay->h_up[(size_t)t * (n - 1)], sizeof(Op) * (n - 1));
  if (down) std::memcpy(down, &lay->h_down[(size_t)t * (n - 1)], sizeof(Op) * (n - 1));
  if (down_nodes) std::memcpy(down_nodes, &lay->h_down_nodes[(size_t)t * (2 * n - 2)], sizeof(int32_t) * (2 * n - 2));
  if (lst) std::memcpy(lst, &lay->h_lst[(size_t)t * N], sizeof(double) * N);
  if (siblings) std::memcpy(siblings, &lay->h_siblings[(size_t)t * N], sizeof(int32_t) * N);
  return VBSKY_OK;
}
