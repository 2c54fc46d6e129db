/*
 * TEST INFRASTRUCTURE — CPU oracle, not product code.
 *
 * Plain-C restatement of the reference's native preprocessing algorithms.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product path (CUDA) never does.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this file is
 * pinned against the UNMODIFIED reference C++ compiled into oracle/_ref/libkpref.so
 * (tests/test_oracle.py) and against fixtures under tests/golden/ generated from it.
 *
 * All paths below are relative to
 *   /root/reference/models/backbone_kpconv/cpp_wrappers/
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "prime_growth.h"

/* ------------------------------------------------------------------------------------------------
 * Voxel-grid barycentre subsampling
 *   follows cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 (one cloud) and :109-211
 *   (stacked batch), with PointXYZ arithmetic from cpp_utils/cloud/cloud.h:54-143.
 *
 * The reference accumulates into std::unordered_map<size_t, SampledData> and emits in the map's
 * ITERATION order, so the container's behaviour is part of the algorithm.  libstdc++'s
 * _Hashtable (unique keys, identity hash, no cached hash codes, max_load_factor 1) is restated:
 *   - one singly linked list through all nodes, headed by a "before begin" sentinel;
 *   - bucket[b] holds the node BEFORE the first node of bucket b (or none);
 *   - insert into an empty bucket: node becomes the global list head;
 *     insert into a non-empty bucket: node goes right after that bucket's "before" node;
 *   - before inserting element n+1 when n+1 > bucket_count: rehash to the next size in
 *     KP_TABLE_SIZES (prime_growth.h), re-inserting the nodes front-to-back with the same rule.
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  uint64_t key;
  int32_t next; /* node index, -1 = end */
  int32_t count;
  float sx, sy, sz;
} vox_node;

typedef struct {
  vox_node* nodes;
  int32_t n_nodes;
  int32_t head;      /* first node of the list (sentinel.next) */
  int32_t* bucket;   /* index of the node before the bucket's first node; -2 = sentinel; -1 = empty */
  uint64_t n_bucket;
  int gen;           /* index into KP_TABLE_SIZES, -1 before first insertion */
} vox_map;

#define BEFORE_BEGIN (-2)

static void map_link_at_bucket_begin(vox_map* m, uint64_t b, int32_t node) {
  if (m->bucket[b] != -1) {
    int32_t before = m->bucket[b];
    int32_t* slot = (before == BEFORE_BEGIN) ? &m->head : &m->nodes[before].next;
    m->nodes[node].next = *slot;
    *slot = node;
  } else {
    m->nodes[node].next = m->head;
    m->head = node;
    if (m->nodes[node].next != -1) {
      uint64_t nb = m->nodes[m->nodes[node].next].key % m->n_bucket;
      m->bucket[nb] = node;
    }
    m->bucket[b] = BEFORE_BEGIN;
  }
}

static int map_rehash(vox_map* m, uint64_t n_bucket) {
  int32_t* nb = (int32_t*)malloc(sizeof(int32_t) * n_bucket);
  if (!nb) return -1;
  for (uint64_t i = 0; i < n_bucket; ++i) nb[i] = -1;
  free(m->bucket);
  m->bucket = nb;
  m->n_bucket = n_bucket;
  int32_t p = m->head;
  m->head = -1;
  /* libstdc++ _M_rehash_aux(unique): walk the old list front to back.  A node landing in an empty
   * bucket becomes the list head; otherwise it is linked right after that bucket's before-node. */
  while (p != -1) {
    int32_t next = m->nodes[p].next;
    map_link_at_bucket_begin(m, m->nodes[p].key % n_bucket, p);
    p = next;
  }
  return 0;
}

static int32_t map_find(const vox_map* m, uint64_t key) {
  if (m->n_bucket == 0) return -1;
  uint64_t b = key % m->n_bucket;
  int32_t before = m->bucket[b];
  if (before == -1) return -1;
  int32_t p = (before == BEFORE_BEGIN) ? m->head : m->nodes[before].next;
  while (p != -1 && m->nodes[p].key % m->n_bucket == b) {
    if (m->nodes[p].key == key) return p;
    p = m->nodes[p].next;
  }
  return -1;
}

static int32_t map_emplace(vox_map* m, uint64_t key) {
  /* _M_need_rehash(n_bkt, n_elt, 1): grows when n_elt + 1 > n_bkt (load factor 1.0) */
  if ((uint64_t)m->n_nodes + 1 > m->n_bucket) {
    m->gen += 1;
    if (m->gen >= KP_N_TABLE_SIZES) return -1;
    if (map_rehash(m, KP_TABLE_SIZES[m->gen]) != 0) return -1;
  }
  int32_t node = m->n_nodes++;
  m->nodes[node].key = key;
  m->nodes[node].count = 0;
  m->nodes[node].sx = m->nodes[node].sy = m->nodes[node].sz = 0.0f;
  map_link_at_bucket_begin(m, key % m->n_bucket, node);
  return node;
}

/* One cloud.  out must hold n*3 floats.  Returns the number of voxels, or <0 on error. */
static int subsample_one(const float* pts, int n, float dl, float* out) {
  if (n <= 0) return 0;
  /* grid_subsampling.cpp:25-27 (bbox), cloud.cpp min_point/max_point */
  float mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      float v = pts[3 * i + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  /* originCorner = floor(minCorner * (1/sampleDl)) * sampleDl, all in fp32 (:27) */
  float inv = 1 / dl;
  float org[3];
  for (int d = 0; d < 3; ++d) org[d] = floorf(mn[d] * inv) * dl;
  /* grid dimensions (:30-31) */
  uint64_t nx = (uint64_t)floorf((mx[0] - org[0]) / dl) + 1;
  uint64_t ny = (uint64_t)floorf((mx[1] - org[1]) / dl) + 1;

  vox_map m;
  m.nodes = (vox_node*)malloc(sizeof(vox_node) * (size_t)n);
  m.n_nodes = 0;
  m.head = -1;
  m.bucket = NULL;
  m.n_bucket = 0;
  m.gen = -1;
  if (!m.nodes) return -1;
  for (int i = 0; i < n; ++i) {
    const float* p = pts + 3 * i;
    /* :53-56 */
    uint64_t ix = (uint64_t)floorf((p[0] - org[0]) / dl);
    uint64_t iy = (uint64_t)floorf((p[1] - org[1]) / dl);
    uint64_t iz = (uint64_t)floorf((p[2] - org[2]) / dl);
    uint64_t key = ix + nx * iy + nx * ny * iz;
    int32_t node = map_find(&m, key);
    if (node < 0) {
      node = map_emplace(&m, key);
      if (node < 0) { free(m.nodes); free(m.bucket); return -1; }
    }
    /* SampledData::update_points (grid_subsampling.h:74-79): sequential fp32 sums */
    m.nodes[node].count += 1;
    m.nodes[node].sx += p[0];
    m.nodes[node].sy += p[1];
    m.nodes[node].sz += p[2];
  }
  /* :85-87: iterate the map, point * (1.0 / count); the double reciprocal narrows to float at the
   * PointXYZ * float operator (cloud.h:120-123). */
  int k = 0;
  for (int32_t p = m.head; p != -1; p = m.nodes[p].next, ++k) {
    float a = (float)(1.0 / (double)m.nodes[p].count);
    out[3 * k + 0] = m.nodes[p].sx * a;
    out[3 * k + 1] = m.nodes[p].sy * a;
    out[3 * k + 2] = m.nodes[p].sz * a;
  }
  free(m.nodes);
  free(m.bucket);
  return k;
}

/* Stacked batch (grid_subsampling.cpp:109-211).  out_pts must hold n*3 floats.  Returns M. */
int kporacle_subsample_batch(const float* pts, int n, const int* lens, int n_clouds, float dl,
                             int max_p, float* out_pts, int* out_lens) {
  if (max_p < 1) max_p = n; /* :134-135 */
  int start = 0, m_total = 0;
  float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  if (!tmp) return -1;
  for (int b = 0; b < n_clouds; ++b) {
    int m = subsample_one(pts + 3 * (size_t)start, lens[b], dl, tmp);
    if (m < 0) { free(tmp); return -1; }
    if (m > max_p) m = max_p; /* :180-205: keep the first max_p */
    memcpy(out_pts + 3 * (size_t)m_total, tmp, sizeof(float) * 3 * (size_t)m);
    out_lens[b] = m;
    m_total += m;
    start += lens[b];
  }
  free(tmp);
  return m_total;
}

/* ------------------------------------------------------------------------------------------------
 * Radius neighbours
 *   follows cpp_neighbors/neighbors/neighbors.cpp:211-332 (batch_nanoflann_neighbors) with the
 *   arithmetic of cpp_utils/nanoflann/nanoflann.hpp:
 *     :432-440  L2_Simple_Adaptor::evalMetric   d2 = ((0 + dx*dx) + dy*dy) + dz*dz in fp32, no FMA
 *     :249-253  RadiusResultSet::addPoint       keep when d2 <  r*r   (strict)
 *     :1280-1289 radiusSearch(sorted=true)      ascending d2
 *   The KD-tree is an accelerator, not part of the result, and is not restated: every support
 *   point of the same cloud is tested.  Equal-d2 ties: the reference's order is whatever an
 *   unstable std::sort leaves; this restatement breaks ties by ascending index.
 *   Row layout (:304-327): local index + cloud offset, padded with the TOTAL support count.
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  float d2;
  int idx;
} hit_t;

static int hit_cmp(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a;
  const hit_t* y = (const hit_t*)b;
  if (x->d2 < y->d2) return -1;
  if (x->d2 > y->d2) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* Returns max_count; *out is malloc'ed [nq * max_count] ints (free with kporacle_free).
 * If tie_rows != NULL it must hold nq bytes and gets 1 where a row contains two equal d2. */
int kporacle_batch_query(const float* q, int nq, const float* s, int ns, const int* q_lens,
                         const int* s_lens, int n_clouds, float radius, int** out,
                         unsigned char* tie_rows) {
  float r2 = radius * radius; /* neighbors.cpp:226 */
  hit_t** rows = (hit_t**)calloc((size_t)(nq > 0 ? nq : 1), sizeof(hit_t*));
  int* counts = (int*)calloc((size_t)(nq > 0 ? nq : 1), sizeof(int));
  hit_t* scratch = (hit_t*)malloc(sizeof(hit_t) * (size_t)(ns > 0 ? ns : 1));
  int max_count = 0, q0 = 0, s0 = 0;
  for (int b = 0; b < n_clouds; ++b) {
    for (int i = q0; i < q0 + q_lens[b]; ++i) {
      int c = 0;
      for (int j = 0; j < s_lens[b]; ++j) {
        const float* sp = s + 3 * (size_t)(s0 + j);
        float dx = q[3 * (size_t)i + 0] - sp[0];
        float dy = q[3 * (size_t)i + 1] - sp[1];
        float dz = q[3 * (size_t)i + 2] - sp[2];
        float d2 = 0.0f;
        d2 += dx * dx;
        d2 += dy * dy;
        d2 += dz * dz;
        if (d2 < r2) { scratch[c].d2 = d2; scratch[c].idx = j + s0; ++c; }
      }
      qsort(scratch, (size_t)c, sizeof(hit_t), hit_cmp);
      if (tie_rows) {
        tie_rows[i] = 0;
        for (int k = 1; k < c; ++k) if (scratch[k].d2 == scratch[k - 1].d2) tie_rows[i] = 1;
      }
      rows[i] = (hit_t*)malloc(sizeof(hit_t) * (size_t)(c > 0 ? c : 1));
      memcpy(rows[i], scratch, sizeof(hit_t) * (size_t)c);
      counts[i] = c;
      if (c > max_count) max_count = c;
    }
    q0 += q_lens[b];
    s0 += s_lens[b];
  }
  int* o = (int*)malloc(sizeof(int) * ((size_t)nq * (size_t)max_count + 1));
  for (int i = 0; i < nq; ++i) {
    for (int k = 0; k < max_count; ++k)
      o[(size_t)i * max_count + k] = k < counts[i] ? rows[i][k].idx : ns; /* pad = supports.size() */
    free(rows[i]);
  }
  free(rows);
  free(counts);
  free(scratch);
  *out = o;
  return max_count;
}

void kporacle_free(void* p) { free(p); }
