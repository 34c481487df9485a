/* TEST INFRASTRUCTURE ONLY - plain-C CPU restatement of the reference's LBG hot path.
 *
 * This file is the executable specification the CUDA path is checked against on the GPU box
 * (where /root/reference does not exist).  It is itself pinned against the real reference
 * (oracle/_ref, built from the unmodified sources) by tests/test_oracle.py and the golden
 * vectors under tests/golden/ that the real reference generated.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product path (quant_b200/) never does.
 *
 * Build: gcc -std=c11 -O2 -fno-fast-math -ffp-contract=off  (strict IEEE double, no FMA) - the
 * semantics of the reference's sources without -ffast-math.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NORMAL 0
#define ORC_SCALED 1
#define ORC_CIE1931 2
#define ORC_LEAF_MAX 10 /* src/KDTree.cpp:4 KD_LEAF_MAX_SIZE; KDTreeVectorOfVectorsAdaptor.hpp:59 */

/* ------------------------------------------------------------------------------------------
 * Colour space and block layout
 * ---------------------------------------------------------------------------------------- */

/* src/ColorSpace.cpp:4-6 (NORMAL) and :16-21 (SCALED); RGB is std::array<char,3>, i.e. the PPM
 * byte reinterpreted as SIGNED (include/RGBImage.hpp:12). */
static double orc_color(uint8_t u, int cs) {
  double s = (double)(int8_t)u;
  if (cs == ORC_SCALED) return (s + 128.0) / 255;
  return s;
}

/* src/ColorSpace.cpp:35-39 Cie1931::RGBtoColorSpace: c[i] are SIGNED chars promoted to double; the sums are
 * evaluated left to right and divided by 0.17697 (plain x86-64 build: no fused multiply-add). */
static void orc_cie_forward(const uint8_t *px, double *out) {
  const double c0 = (double)(int8_t)px[0], c1 = (double)(int8_t)px[1], c2 = (double)(int8_t)px[2];
  out[0] = (c0 * 0.490 + c1 * 0.310 + c2 * 0.200) / 0.17697;
  out[1] = (c0 * 0.17697 + c1 * 0.81240 + c2 * 0.01063) / 0.17697;
  out[2] = (c0 * 0 + c1 * 0.01 + c2 * 0.99) / 0.17697;
}

/* Integer lattice value of a byte: SCALED value == t/255.0 with t = u ^ 0x80 in [0,255];
 * NORMAL value == (int8)u.   */
int orc_lattice(uint8_t u, int cs) { return cs == ORC_SCALED ? (int)(u ^ 0x80) : (int)(int8_t)u; }

size_t orc_num_vectors(int xSize, int ySize, int w, int h) {
  size_t wB = ((size_t)xSize + w - 1) / w, hB = ((size_t)ySize + h - 1) / h;
  return wB * hB;
}

/* src/Compressor.cpp:31-62 getBlocksAsVectorsFromImage.  The buffer is indexed x*ySize + y
 * (:49) - x is the SLOW axis - and only imgIndex >= xSize*ySize yields 0 (:53-57); a block that
 * overflows in y wraps into the next x line. out: N*dim doubles, vector i*hBlocks + j (:59). */
void orc_blocks(const uint8_t *rgb, int xSize, int ySize, int w, int h, int cs, double *out) {
  size_t wB = ((size_t)xSize + w - 1) / w, hB = ((size_t)ySize + h - 1) / h;
  size_t npix = (size_t)xSize * ySize, dim = 3 * (size_t)w * h;
  for (size_t i = 0; i < wB; i++)
    for (size_t j = 0; j < hB; j++) {
      double *v = out + (i * hB + j) * dim;
      for (size_t x = i * w; x < i * w + w; x++)
        for (size_t y = j * h; y < j * h + h; y++) {
          size_t img = x * ySize + y;
          size_t e = ((x - i * w) * h + (y - j * h)) * 3;
          if (cs == ORC_CIE1931 && img < npix)
            orc_cie_forward(rgb + img * 3, v + e);
          else
            for (int ch = 0; ch < 3; ch++) v[e + ch] = img < npix ? orc_color(rgb[img * 3 + ch], cs) : 0.0;
        }
    }
}

/* Same walk, emitting lattice integers (int16) plus a validity flag per element: elements past
 * the buffer end are the colour-space value 0.0, which is lattice 0 in both spaces. */
void orc_blocks_lattice(const uint8_t *rgb, int xSize, int ySize, int w, int h, int cs, int16_t *out) {
  size_t wB = ((size_t)xSize + w - 1) / w, hB = ((size_t)ySize + h - 1) / h;
  size_t npix = (size_t)xSize * ySize, dim = 3 * (size_t)w * h;
  for (size_t i = 0; i < wB; i++)
    for (size_t j = 0; j < hB; j++) {
      int16_t *v = out + (i * hB + j) * dim;
      for (size_t x = i * w; x < i * w + w; x++)
        for (size_t y = j * h; y < j * h + h; y++) {
          size_t img = x * ySize + y;
          size_t e = ((x - i * w) * h + (y - j * h)) * 3;
          for (int ch = 0; ch < 3; ch++)
            v[e + ch] = img < npix ? (int16_t)orc_lattice(rgb[img * 3 + ch], cs) : 0;
        }
    }
}

/* src/Compressor.cpp:12-29 + src/ColorSpace.cpp:8-11,23-28: codebook doubles -> bytes.
 * SCALED: (char)std::round((c - 128.0) * 255) - only right through int8 wrap-around; the x86
 * build converts through a 32-bit integer, stated explicitly here. */
void orc_codebook_to_bytes(const double *cb, size_t K, int dim, int cs, uint8_t *out) {
  if (cs == ORC_CIE1931) { /* src/ColorSpace.cpp:41-48 Cie1931::colorSpaceToRGB, one pixel (3 elements) at a time */
    for (size_t i = 0; i + 2 < K * (size_t)dim; i += 3) {
      const double *c = cb + i;
      double t0 = (c[0] * 0.418 + c[1] * (-0.15866) + c[2] * (-0.082835));
      double t1 = (c[0] * (-0.091169) + c[1] * 0.25243 + c[2] * 0.015708);
      double t2 = (c[0] * 0.0009209 + c[1] * (-0.0025498) + c[2] * 0.17860);
      out[i] = (uint8_t)(int8_t)(int)round(t0);
      out[i + 1] = (uint8_t)(int8_t)(int)round(t1);
      out[i + 2] = (uint8_t)(int8_t)(int)round(t2);
    }
    return;
  }
  for (size_t i = 0; i < K * (size_t)dim; i++) {
    double r = cs == ORC_SCALED ? round((cb[i] - 128.0) * 255) : round(cb[i]);
    out[i] = (uint8_t)(int8_t)(int)r;
  }
}

/* src/Compressor.cpp:156-165 decompress + :64-92 getImageFromVectors.  Pixels no block writes
 * keep RGB{} == 0 (value-initialised std::array). */
void orc_decode(const uint8_t *cb_bytes, const uint64_t *assign, int xSize, int ySize, int w, int h,
                uint8_t *rgb_out) {
  size_t wB = ((size_t)xSize + w - 1) / w, hB = ((size_t)ySize + h - 1) / h;
  size_t npix = (size_t)xSize * ySize, dim = 3 * (size_t)w * h;
  memset(rgb_out, 0, npix * 3);
  for (size_t i = 0; i < wB; i++)
    for (size_t j = 0; j < hB; j++) {
      const uint8_t *v = cb_bytes + assign[i * hB + j] * dim;
      for (size_t x = i * w; x < i * w + w; x++)
        for (size_t y = j * h; y < j * h + h; y++) {
          size_t img = x * ySize + y;
          size_t e = ((x - i * w) * h + (y - j * h)) * 3;
          if (img < npix)
            for (int ch = 0; ch < 3; ch++) rgb_out[img * 3 + ch] = v[e + ch];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * nanoflann 1.2.3 KD-tree, restated (include/external/nanoflann.hpp)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  double low, high;
} orc_interval;

typedef struct {
  int32_t child1, child2; /* -1/-1 => leaf */
  int32_t divfeat;
  uint32_t left, right; /* leaf: range in vind */
  double divlow, divhigh;
} orc_node;

typedef struct orc_kdtree {
  const double *pts;
  size_t n;
  int dim;
  size_t *vind;
  orc_node *nodes;
  size_t n_nodes, cap_nodes;
  orc_interval *root_bbox;
  int max_depth;
} orc_kdtree;

#define PT(t, idx, d) ((t)->pts[(idx) * (size_t)(t)->dim + (d)])

static int32_t orc_new_node(orc_kdtree *t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 64;
    t->nodes = (orc_node *)realloc(t->nodes, t->cap_nodes * sizeof(orc_node));
  }
  return (int32_t)t->n_nodes++;
}

/* nanoflann.hpp:1096-1105 computeMinMax */
static void orc_minmax(const orc_kdtree *t, const size_t *ind, size_t count, int el, double *mn, double *mx) {
  *mn = *mx = PT(t, ind[0], el);
  for (size_t i = 1; i < count; i++) {
    double v = PT(t, ind[i], el);
    if (v < *mn) *mn = v;
    if (v > *mx) *mx = v;
  }
}

/* nanoflann.hpp:1159-1186 planeSplit (IndexType = size_t, hence the "!right" guards) */
static void orc_plane_split(const orc_kdtree *t, size_t *ind, size_t count, int cutfeat, double cutval,
                            size_t *lim1, size_t *lim2) {
  size_t left = 0, right = count - 1;
  for (;;) {
    while (left <= right && PT(t, ind[left], cutfeat) < cutval) ++left;
    while (right && left <= right && PT(t, ind[right], cutfeat) >= cutval) --right;
    if (left > right || !right) break;
    size_t tmp = ind[left];
    ind[left] = ind[right];
    ind[right] = tmp;
    ++left;
    --right;
  }
  *lim1 = left;
  right = count - 1;
  for (;;) {
    while (left <= right && PT(t, ind[left], cutfeat) <= cutval) ++left;
    while (right && left <= right && PT(t, ind[right], cutfeat) > cutval) --right;
    if (left > right || !right) break;
    size_t tmp = ind[left];
    ind[left] = ind[right];
    ind[right] = tmp;
    ++left;
    --right;
  }
  *lim2 = left;
}

/* nanoflann.hpp:1108-1147 middleSplit_ */
static void orc_middle_split(const orc_kdtree *t, size_t *ind, size_t count, size_t *index, int *cutfeat,
                             double *cutval, const orc_interval *bbox) {
  const double EPS = (double)0.00001;
  int dim = t->dim;
  double max_span = bbox[0].high - bbox[0].low;
  for (int i = 1; i < dim; ++i) {
    double span = bbox[i].high - bbox[i].low;
    if (span > max_span) max_span = span;
  }
  double max_spread = -1;
  *cutfeat = 0;
  for (int i = 0; i < dim; ++i) {
    double span = bbox[i].high - bbox[i].low;
    if (span > (1 - EPS) * max_span) {
      double mn, mx;
      orc_minmax(t, ind, count, i, &mn, &mx);
      double spread = mx - mn;
      if (spread > max_spread) {
        *cutfeat = i;
        max_spread = spread;
      }
    }
  }
  double split_val = (bbox[*cutfeat].low + bbox[*cutfeat].high) / 2;
  double mn, mx;
  orc_minmax(t, ind, count, *cutfeat, &mn, &mx);
  if (split_val < mn)
    *cutval = mn;
  else if (split_val > mx)
    *cutval = mx;
  else
    *cutval = split_val;
  size_t lim1, lim2;
  orc_plane_split(t, ind, count, *cutfeat, *cutval, &lim1, &lim2);
  if (lim1 > count / 2)
    *index = lim1;
  else if (lim2 < count / 2)
    *index = lim2;
  else
    *index = count / 2;
}

/* nanoflann.hpp:1046-1094 divideTree */
static int32_t orc_divide(orc_kdtree *t, size_t left, size_t right, orc_interval *bbox, int depth) {
  int dim = t->dim;
  int32_t me = orc_new_node(t);
  if (depth > t->max_depth) t->max_depth = depth;
  if ((right - left) <= (size_t)ORC_LEAF_MAX) {
    t->nodes[me].child1 = t->nodes[me].child2 = -1;
    t->nodes[me].left = (uint32_t)left;
    t->nodes[me].right = (uint32_t)right;
    t->nodes[me].divfeat = 0;
    t->nodes[me].divlow = t->nodes[me].divhigh = 0;
    for (int i = 0; i < dim; ++i) bbox[i].low = bbox[i].high = PT(t, t->vind[left], i);
    for (size_t k = left + 1; k < right; ++k)
      for (int i = 0; i < dim; ++i) {
        double v = PT(t, t->vind[k], i);
        if (bbox[i].low > v) bbox[i].low = v;
        if (bbox[i].high < v) bbox[i].high = v;
      }
  } else {
    size_t idx;
    int cutfeat;
    double cutval;
    orc_middle_split(t, t->vind + left, right - left, &idx, &cutfeat, &cutval, bbox);
    orc_interval *lb = (orc_interval *)malloc(sizeof(orc_interval) * dim);
    orc_interval *rb = (orc_interval *)malloc(sizeof(orc_interval) * dim);
    memcpy(lb, bbox, sizeof(orc_interval) * dim);
    lb[cutfeat].high = cutval;
    int32_t c1 = orc_divide(t, left, left + idx, lb, depth + 1);
    memcpy(rb, bbox, sizeof(orc_interval) * dim);
    rb[cutfeat].low = cutval;
    int32_t c2 = orc_divide(t, left + idx, right, rb, depth + 1);
    orc_node *nd = &t->nodes[me]; /* re-fetch: realloc may have moved the pool */
    nd->child1 = c1;
    nd->child2 = c2;
    nd->divfeat = cutfeat;
    nd->left = nd->right = 0;
    nd->divlow = lb[cutfeat].high;
    nd->divhigh = rb[cutfeat].low;
    for (int i = 0; i < dim; ++i) {
      bbox[i].low = rb[i].low < lb[i].low ? rb[i].low : lb[i].low;       /* std::min(l, r) */
      bbox[i].high = lb[i].high < rb[i].high ? rb[i].high : lb[i].high; /* std::max(l, r) */
    }
    free(lb);
    free(rb);
  }
  return me;
}

/* nanoflann.hpp:863-871 buildIndex + :1021-1043 computeBoundingBox.  (The reference builds the
 * index twice - adaptor ctor and src/KDTree.cpp:11 - with identical results.) */
orc_kdtree *orc_kd_build(const double *pts, size_t n, int dim) {
  orc_kdtree *t = (orc_kdtree *)calloc(1, sizeof(orc_kdtree));
  t->pts = pts;
  t->n = n;
  t->dim = dim;
  t->vind = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) t->vind[i] = i;
  t->root_bbox = (orc_interval *)malloc(sizeof(orc_interval) * dim);
  if (n == 0) return t;
  for (int i = 0; i < dim; ++i) t->root_bbox[i].low = t->root_bbox[i].high = PT(t, 0, i);
  for (size_t k = 1; k < n; ++k)
    for (int i = 0; i < dim; ++i) {
      double v = PT(t, k, i);
      if (v < t->root_bbox[i].low) t->root_bbox[i].low = v;
      if (v > t->root_bbox[i].high) t->root_bbox[i].high = v;
    }
  orc_divide(t, 0, n, t->root_bbox, 1);
  return t;
}

void orc_kd_free(orc_kdtree *t) {
  if (!t) return;
  free(t->vind);
  free(t->nodes);
  free(t->root_bbox);
  free(t);
}

int orc_kd_depth(const orc_kdtree *t) { return t->max_depth; }
void orc_kd_order(const orc_kdtree *t, uint32_t *out) {
  for (size_t i = 0; i < t->n; i++) out[i] = (uint32_t)t->vind[i];
}
size_t orc_kd_num_nodes(const orc_kdtree *t) { return t->n_nodes; }

/* nanoflann.hpp:320-345 L2_Adaptor::operator() with worst_dist = -1 (no early exit): groups of
 * four, ((d0^2 + d1^2) + d2^2) + d3^2 added to the running sum, then a scalar tail. */
double orc_l2(const double *a, const double *b, int size) {
  double result = 0.0;
  int d = 0;
  for (; d + 3 < size; d += 4) {
    const double diff0 = a[d] - b[d];
    const double diff1 = a[d + 1] - b[d + 1];
    const double diff2 = a[d + 2] - b[d + 2];
    const double diff3 = a[d + 3] - b[d + 3];
    result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
  }
  for (; d < size; d++) {
    const double diff0 = a[d] - b[d];
    result += diff0 * diff0;
  }
  return result;
}

typedef struct {
  double best;
  size_t idx;
  size_t count;
} orc_result; /* KNNResultSet, capacity 1 (nanoflann.hpp:78-144) */

/* nanoflann.hpp:1213-1270 searchLevel (epsError == 1.0f) */
static void orc_search(const orc_kdtree *t, orc_result *rs, const double *vec, int32_t ni, double mindistsq,
                       double *dists) {
  const orc_node *node = &t->nodes[ni];
  if (node->child1 < 0 && node->child2 < 0) {
    double worst = rs->best; /* read once per leaf (:1219) */
    for (size_t i = node->left; i < node->right; ++i) {
      size_t index = t->vind[i];
      double dist = orc_l2(vec, t->pts + index * (size_t)t->dim, t->dim);
      if (dist < worst) {
        /* addPoint (:114-138) with capacity 1: replaces only when dists[0] > dist (strict) */
        if (rs->count == 0 || rs->best > dist) {
          rs->best = dist;
          rs->idx = index;
        }
        rs->count = 1;
      }
    }
    return;
  }
  int idx = node->divfeat;
  double val = vec[idx];
  double diff1 = val - node->divlow;
  double diff2 = val - node->divhigh;
  int32_t bestChild, otherChild;
  double cut_dist;
  if ((diff1 + diff2) < 0) {
    bestChild = node->child1;
    otherChild = node->child2;
    cut_dist = (val - node->divhigh) * (val - node->divhigh);
  } else {
    bestChild = node->child2;
    otherChild = node->child1;
    cut_dist = (val - node->divlow) * (val - node->divlow);
  }
  orc_search(t, rs, vec, bestChild, mindistsq, dists);
  double dst = dists[idx];
  mindistsq = mindistsq + cut_dist - dst;
  dists[idx] = cut_dist;
  if (mindistsq * (double)1.0f <= rs->best) orc_search(t, rs, vec, otherChild, mindistsq, dists);
  dists[idx] = dst;
}

/* nanoflann.hpp:906-920 findNeighbors + :1188-1205 computeInitialDistances;
 * src/KDTree.cpp:20-29 nearestNeighbour. */
size_t orc_kd_nn(const orc_kdtree *t, const double *vec) {
  double dists_stack[64];
  double *dists = t->dim <= 64 ? dists_stack : (double *)malloc(sizeof(double) * t->dim);
  double distsq = 0.0;
  for (int i = 0; i < t->dim; ++i) {
    dists[i] = 0;
    if (vec[i] < t->root_bbox[i].low) {
      dists[i] = (vec[i] - t->root_bbox[i].low) * (vec[i] - t->root_bbox[i].low);
      distsq += dists[i];
    }
    if (vec[i] > t->root_bbox[i].high) {
      dists[i] = (vec[i] - t->root_bbox[i].high) * (vec[i] - t->root_bbox[i].high);
      distsq += dists[i];
    }
  }
  orc_result rs;
  rs.best = 1.7976931348623157e308; /* numeric_limits<double>::max() (:97) */
  rs.idx = 0;
  rs.count = 0;
  orc_search(t, &rs, vec, 0, distsq, dists);
  if (dists != dists_stack) free(dists);
  return rs.idx;
}

/* src/Quantizer.cpp:24-32 assignCodeVectors */
void orc_assign(const double *X, size_t N, int dim, const double *cb, size_t K, uint64_t *assign) {
  orc_kdtree *t = orc_kd_build(cb, K, dim);
  for (size_t i = 0; i < N; i++) assign[i] = orc_kd_nn(t, X + i * (size_t)dim);
  orc_kd_free(t);
}

/* Exhaustive first-minimum search with the same distance arithmetic.  NOT what the reference
 * does at exact ties - kept to measure how often the KD traversal order matters. */
void orc_assign_bruteforce(const double *X, size_t N, int dim, const double *cb, size_t K, uint64_t *assign) {
  for (size_t i = 0; i < N; i++) {
    double best = 1.7976931348623157e308;
    size_t bi = 0;
    for (size_t k = 0; k < K; k++) {
      double d = orc_l2(X + i * (size_t)dim, cb + k * (size_t)dim, dim);
      if (d < best) {
        best = d;
        bi = k;
      }
    }
    assign[i] = bi;
  }
}

/* ------------------------------------------------------------------------------------------
 * Quantizer.cpp restated
 * ---------------------------------------------------------------------------------------- */

/* src/Quantizer.cpp:9-22 updateDistortion; norm = left-to-right sum of squares
 * (include/VectorOperations.hpp:107-111).  The reference's OpenMP reduction leaves the
 * cross-thread order unspecified; this is its single-thread order. */
double orc_distortion(const double *X, size_t N, int dim, const double *cb, const uint64_t *assign) {
  double res = 0;
  for (size_t i = 0; i < N; i++) {
    const double *x = X + i * (size_t)dim, *c = cb + assign[i] * (size_t)dim;
    double nrm = 0.0;
    for (int d = 0; d < dim; d++) {
      double v = x[d] - c[d];
      nrm += v * v;
    }
    res += nrm;
  }
  return res / ((double)(N * (size_t)dim));
}

/* src/Quantizer.cpp:46-57 trainingSetSum: element-wise Kahan sum in index order. */
void orc_training_sum(const double *X, size_t N, int dim, double *sum) {
  double *c = (double *)calloc(dim, sizeof(double));
  for (int d = 0; d < dim; d++) sum[d] = 0;
  for (size_t i = 0; i < N; i++)
    for (int d = 0; d < dim; d++) {
      double y = X[i * (size_t)dim + d] - c[d];
      double t = sum[d] + y;
      c[d] = (t - sum[d]) - y;
      sum[d] = t;
    }
  free(c);
}

/* src/Quantizer.cpp:72-87 fixCodeVectors + :59-70 sumInArea: per cell, Kahan sum of members in
 * ascending index order, divided by the member count; an empty cell becomes the zero vector. */
void orc_fix(const double *X, size_t N, int dim, const uint64_t *assign, size_t K, double *cb) {
  double *c = (double *)calloc(K * (size_t)dim, sizeof(double));
  size_t *cnt = (size_t *)calloc(K, sizeof(size_t));
  for (size_t i = 0; i < K * (size_t)dim; i++) cb[i] = 0;
  for (size_t i = 0; i < N; i++) {
    size_t k = assign[i];
    cnt[k]++;
    for (int d = 0; d < dim; d++) {
      double *s = &cb[k * (size_t)dim + d], *cc = &c[k * (size_t)dim + d];
      double y = X[i * (size_t)dim + d] - *cc;
      double t = *s + y;
      *cc = (t - *s) - y;
      *s = t;
    }
  }
  for (size_t k = 0; k < K; k++)
    if (cnt[k])
      for (int d = 0; d < dim; d++) cb[k * (size_t)dim + d] /= (double)cnt[k];
  free(c);
  free(cnt);
}

/* src/Quantizer.cpp:134-138: concat(cb, cb); first half *= (1 + 0.2), second half *= (1 - 0.2).
 * The factors are the doubles those C expressions evaluate to (NOT the literals 1.2 / 0.8:
 * 1 - 0.2 rounds to 0.79999999999999993 < 0.8). cb holds K entries in, 2K out. */
void orc_split(double *cb, size_t K, int dim) {
  for (size_t i = 0; i < K * (size_t)dim; i++) {
    double v = cb[i];
    cb[i] = v * (double)(1 + 0.2);
    cb[K * (size_t)dim + i] = v * (double)(1 - 0.2);
  }
}

/* src/Quantizer.cpp:121-143 LBGQuantizer::quantize with :98-108 LBGIterate stated literally
 * (one assignment per level; fix/distortion loop until |old-new|/old <= eps, max 100 rounds).
 * Optional per-level dumps as in oracle/ref_harness.cpp::ref_levels. Returns K. */
size_t orc_quantize(const double *X, size_t N, int dim, int nbits, double eps, double *cb_out,
                    uint64_t *assign_out, double *dist_out, double *cb0, double *cb_pre, uint64_t *assign_lv,
                    double *cb_post, double *d0, double *d1, int *iters) {
  size_t maxK = (size_t)1 << nbits;
  double *cb = (double *)malloc(sizeof(double) * maxK * dim);
  orc_training_sum(X, N, dim, cb);
  for (int d = 0; d < dim; d++) cb[d] /= (double)N;
  if (cb0) memcpy(cb0, cb, sizeof(double) * dim);
  size_t K = 1, off = 0;
  double distortion = 0;
  int level = 0;
  while (K < maxK) {
    orc_split(cb, K, dim);
    K *= 2;
    if (cb_pre) memcpy(cb_pre + off, cb, sizeof(double) * K * dim);
    orc_assign(X, N, dim, cb, K, assign_out);
    if (assign_lv) memcpy(assign_lv + (size_t)level * N, assign_out, sizeof(uint64_t) * N);
    distortion = orc_distortion(X, N, dim, cb, assign_out);
    if (d0) d0[level] = distortion;
    int it;
    for (it = 0; it < 100; it++) {
      orc_fix(X, N, dim, assign_out, K, cb);
      double old = distortion;
      distortion = orc_distortion(X, N, dim, cb, assign_out);
      if (fabs(old - distortion) / old <= eps) {
        it++;
        break;
      }
    }
    if (iters) iters[level] = it;
    if (d1) d1[level] = distortion;
    if (cb_post) memcpy(cb_post + off, cb, sizeof(double) * K * dim);
    off += K * dim;
    level++;
  }
  memcpy(cb_out, cb, sizeof(double) * K * dim);
  *dist_out = distortion;
  free(cb);
  return K;
}

/* ------------------------------------------------------------------------------------------
 * Integer statistics (SURVEY.md section 8a "key simplification"): what the CUDA accumulate kernel
 * produces, and the O(K*dim) FP64 finalisation the product derives from them.
 * ---------------------------------------------------------------------------------------- */

/* n_k, S_k[d] = sum of lattice values, Q_k = sum over members and dims of lattice^2. */
void orc_stats(const int16_t *T, size_t N, int dim, const uint64_t *assign, size_t K, uint64_t *n, int64_t *S,
               uint64_t *Q) {
  memset(n, 0, sizeof(uint64_t) * K);
  memset(S, 0, sizeof(int64_t) * K * dim);
  memset(Q, 0, sizeof(uint64_t) * K);
  for (size_t i = 0; i < N; i++) {
    size_t k = assign[i];
    n[k]++;
    for (int d = 0; d < dim; d++) {
      int v = T[i * (size_t)dim + d];
      S[k * (size_t)dim + d] += v;
      Q[k] += (uint64_t)(v * v);
    }
  }
}

/* Centroids from integer sums: SCALED (S/255.0)/n, NORMAL S/n; empty cell -> zero vector.
 * Differs from orc_fix (Kahan over rounded t/255.0 terms) by at most a few ulp. */
void orc_centroids_from_stats(const uint64_t *n, const int64_t *S, size_t K, int dim, int cs, double *cb) {
  for (size_t k = 0; k < K; k++)
    for (int d = 0; d < dim; d++) {
      double s = (double)S[k * (size_t)dim + d];
      if (cs == ORC_SCALED) s = s / 255.0;
      cb[k * (size_t)dim + d] = n[k] ? s / (double)n[k] : 0.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * .quant container (src/Compressor.cpp:190-227 saveToFile)
 * ---------------------------------------------------------------------------------------- */

/* src/Compressor.cpp:167-172 smallestPow2 */
static size_t orc_log2(size_t n) {
  size_t p = 0;
  while (n /= 2) p++;
  return p;
}

/* Header "<bits> <colorSpace> <N> <xSize> <ySize> <bw> <bh>\n", K*dim codebook bytes, then N
 * indices as the low ceil(bits/8) bytes of a little-endian size_t.  Returns bytes written. */
size_t orc_quant_serialize(const uint8_t *cb_bytes, size_t K, const uint64_t *assign, size_t N, int xSize,
                           int ySize, int w, int h, int colorspace_field, uint8_t *out, size_t cap) {
  size_t bits = orc_log2(K), dim = 3 * (size_t)w * h;
  char hdr[256];
  int hl = snprintf(hdr, sizeof hdr, "%zu %d %zu %d %d %d %d\n", bits, colorspace_field, N, xSize, ySize, w, h);
  size_t bpi = ((bits + 7) / 8);
  size_t total = (size_t)hl + K * dim + N * bpi;
  if (!out) return total;
  if (total > cap) return 0;
  memcpy(out, hdr, hl);
  memcpy(out + hl, cb_bytes, K * dim);
  uint8_t *p = out + hl + K * dim;
  for (size_t i = 0; i < N; i++)
    for (size_t b = 0; b < bpi; b++) *p++ = (uint8_t)(assign[i] >> (8 * b));
  return total;
}
