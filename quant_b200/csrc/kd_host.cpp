// See kd_host.hpp.  Compiled as plain C++ with strict IEEE double (no -ffast-math, no FMA
// contraction: all comparisons and the midpoint below must round exactly like the reference's).
#include "kd_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>

namespace qb {
namespace {

struct Range {
  double lo, hi;
  // robustness census only: the bound is a number the integer-sum and the compensated-sum codebooks share bit for bit
  // (it comes from bit-reproducible points, or from a plane computed from such bounds)
  bool lo_exact = false, hi_exact = false;
};

class Builder {
 public:
  Builder(const double *pts, size_t n, int dim, int leaf_max, KdHostTree &out, const unsigned char *exact)
      : pts_(pts), n_(n), dim_(dim), leaf_max_((size_t)leaf_max), out_(out), exact_(exact) {}

  void run() {
    out_.nodes.clear();
    out_.nodes.reserve(2 * n_ / (leaf_max_ ? leaf_max_ : 1) + 8);
    out_.order.resize(n_);
    for (size_t i = 0; i < n_; i++) out_.order[i] = (unsigned int)i;
    out_.depth = 0;
    out_.min_margin = 1.0;
    std::vector<Range> box(dim_);
    if (n_ == 0) {
      out_.box_low.assign(dim_, 0.0);
      out_.box_high.assign(dim_, 0.0);
      return;
    }
    // computeBoundingBox: strict < / > updates starting from point 0.
    for (int d = 0; d < dim_; d++) box[d].lo = box[d].hi = at(0, d);
    for (size_t k = 1; k < n_; k++)
      for (int d = 0; d < dim_; d++) {
        double v = at(k, d);
        if (v < box[d].lo) box[d].lo = v;
        if (v > box[d].hi) box[d].hi = v;
      }
    for (int d = 0; d < dim_; d++) {
      double mn, mx;
      span_of(0, n_, d, mn, mx, &box[d].lo_exact, &box[d].hi_exact);
    }
    subdivide(0, n_, box, 1);
    out_.box_low.resize(dim_);
    out_.box_high.resize(dim_);
    for (int d = 0; d < dim_; d++) {
      out_.box_low[d] = box[d].lo;
      out_.box_high[d] = box[d].hi;
    }
  }

 private:
  double at(size_t point, int d) const { return pts_[point * (size_t)dim_ + d]; }
  // Robustness census (KdHostTree::min_margin): the smallest relative distance of any comparison that shapes the tree
  // from flipping.  Not part of nanoflann; it only observes.
  void note(double a, double b) {
    // (0 against 0: a subtree of dead cells' zero vectors - the same zeros with either centroid arithmetic)
    if (a == 0.0 && b == 0.0) return;
    const double scale = std::max(std::fabs(a), std::fabs(b));
    const double m = scale > 0 ? std::fabs(a - b) / scale : 0.0;
    if (m < out_.min_margin) out_.min_margin = m;
    if (m < node_margin_) node_margin_ = m;
  }
  void no_margin() {  // a comparison that is a coin toss between the two centroid arithmetics
    out_.min_margin = 0.0;
    node_margin_ = 0.0;
  }
  double via(size_t pos, int d) const { return at(out_.order[pos], d); }

  bool point_exact(size_t pos) const { return exact_ && exact_[out_.order[pos]]; }
  // mn_exact / mx_exact (census only): every point attaining the extreme is bit-reproducible and no other kind of
  // point is within rounding noise of it - the extreme is then the same number with either centroid arithmetic
  void span_of(size_t first, size_t count, int d, double &mn, double &mx, bool *mn_exact = nullptr,
               bool *mx_exact = nullptr) const {
    mn = mx = via(first, d);
    for (size_t i = 1; i < count; i++) {
      double v = via(first + i, d);
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    if (!mn_exact) return;
    *mn_exact = *mx_exact = exact_ != nullptr;
    const double tol = 1e-9 * std::max(std::fabs(mn), std::fabs(mx));
    for (size_t i = 0; i < count && exact_; i++) {
      if (point_exact(first + i)) continue;
      const double v = via(first + i, d);
      if (v <= mn + tol) *mn_exact = false;
      if (v >= mx - tol) *mx_exact = false;
    }
  }

  // planeSplit on order[first .. first+count): two Hoare-style sweeps with unsigned indices
  // (the "!right" exits are part of the behaviour being reproduced).
  void partition(size_t first, size_t count, int feat, double cut, size_t &lim1, size_t &lim2) {
    unsigned int *ind = out_.order.data() + first;
    auto val = [&](size_t p) { return at(ind[p], feat); };
    size_t left = 0, right = count - 1;
    for (;;) {
      while (left <= right && val(left) < cut) ++left;
      while (right && left <= right && val(right) >= cut) --right;
      if (left > right || !right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && val(left) <= cut) ++left;
      while (right && left <= right && val(right) > cut) --right;
      if (left > right || !right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    lim2 = left;
  }

  // middleSplit_: cut the dimension of (nearly) largest box span with the largest actual spread
  // at the box midpoint clamped into the data range; balance by count when the plane is lopsided.
  size_t choose_split(size_t first, size_t count, const std::vector<Range> &box, int &feat,
                      double &cut) {
    const double kEps = static_cast<double>(0.00001);
    node_margin_ = 1.0;  // the smallest margin of the comparisons made for THIS node (subdivide turns it into a census bit)
    double max_span = box[0].hi - box[0].lo;
    for (int d = 1; d < dim_; d++) {
      double span = box[d].hi - box[d].lo;
      if (span > max_span) max_span = span;
    }
    double best_spread = -1;
    bool best_exact = false;
    feat = 0;
    spreads_.clear();
    for (int d = 0; d < dim_; d++) {
      double span = box[d].hi - box[d].lo;
      note(span, (1 - kEps) * max_span);  // eligibility of this dimension
      if (span > (1 - kEps) * max_span) {
        double mn, mx;
        bool mn_e, mx_e;
        span_of(first, count, d, mn, mx, &mn_e, &mx_e);
        double spread = mx - mn;
        spreads_.push_back({spread, d, mn_e && mx_e});
        if (spread > best_spread) {
          feat = d;
          best_spread = spread;
          best_exact = mn_e && mx_e;
        }
      }
    }
    // which eligible dimension wins: the winner against EVERY other eligible one (several dimensions can share the
    // runner-up spread).  A tie between two spreads that are the same numbers in both codebooks is harmless.
    for (const Spread &sp : spreads_)
      if (sp.dim != feat && !(best_exact && sp.exact)) note(best_spread, sp.value);
    double mid = (box[feat].lo + box[feat].hi) / 2;
    const bool mid_exact = box[feat].lo_exact && box[feat].hi_exact;
    double mn, mx;
    bool mn_exact, mx_exact;
    span_of(first, count, feat, mn, mx, &mn_exact, &mx_exact);
    cut = mid < mn ? mn : (mid > mx ? mx : mid);
    cut_exact_ = mid < mn ? mn_exact : (mid > mx ? mx_exact : mid_exact);
    {
      // every point against the cutting plane.  A point AT the plane is harmless when the plane was clamped onto
      // the data range (then the plane is that point's own coordinate and moves with it), or when point and plane
      // are both bit-reproducible numbers.
      const bool clamped = mid < mn || mid > mx;  // (not "cut == mn || cut == mx": an unclamped midpoint can coincide with an extreme)
      if (mn == mx && mn != 0.0 && !(mn_exact && mx_exact)) no_margin();  // all points equal along the cut
      size_t at_cut = 0, at_cut_inexact = 0;
      for (size_t i = 0; i < count; i++) {
        const double v = via(first + i, feat);
        if (v == cut) {
          at_cut++;
          at_cut_inexact += !point_exact(first + i);
        }
        if (cut_exact_ && point_exact(first + i)) continue;
        if (v != cut)
          note(v, cut);
        else if (!clamped)
          no_margin();
      }
      // clamped plane shared by several points, not all of them bit-reproducible: in the other codebook they may sit an
      // ulp apart, and how many are <= the plane (planeSplit's second limit) changes
      if (clamped && at_cut > 1 && at_cut_inexact > 0 && cut != 0.0) no_margin();
      if (!(mid_exact && mn_exact && mx_exact)) {  // the clamp decisions themselves (far from flipping when clamped by a lot)
        note(mid, mn);
        note(mid, mx);
      }
    }
    size_t lim1, lim2;
    partition(first, count, feat, cut, lim1, lim2);
    static const bool trace = std::getenv("QB200_KD_TRACE") != nullptr;  // diagnostics: one line per inner node
    if (trace)
      std::fprintf(stderr, "kd node first=%zu count=%zu feat=%d cut=%.17g (exact %d) lim1=%zu lim2=%zu margin=%.3g  box [%.17g (%d), %.17g (%d)] data [%.17g (%d), %.17g (%d)]\n",
                   first, count, feat, cut, (int)cut_exact_, lim1, lim2, out_.min_margin, box[feat].lo, (int)box[feat].lo_exact, box[feat].hi,
                   (int)box[feat].hi_exact, mn, (int)mn_exact, mx, (int)mx_exact);
    if (lim1 > count / 2) return lim1;
    if (lim2 < count / 2) return lim2;
    return count / 2;
  }

  // divideTree: `box` is in/out - on return it is the tight box of the subtree.
  int subdivide(size_t first, size_t last, std::vector<Range> &box, int level) {
    int me = (int)out_.nodes.size();
    out_.nodes.emplace_back();
    if (level > out_.depth) out_.depth = level;
    if (last - first <= leaf_max_) {
      KdNode &nd = out_.nodes[me];
      nd.child1 = nd.child2 = -1;
      nd.a = (int)first;
      nd.b = (int)last;
      nd.divlow = nd.divhigh = 0.0;
      for (int d = 0; d < dim_; d++) box[d].lo = box[d].hi = via(first, d);
      for (size_t p = first + 1; p < last; p++)
        for (int d = 0; d < dim_; d++) {
          double v = via(p, d);
          if (box[d].lo > v) box[d].lo = v;
          if (box[d].hi < v) box[d].hi = v;
        }
      return me;
    }
    int feat;
    double cut;
    size_t nleft = choose_split(first, last - first, box, feat, cut);
    const bool cut_exact = cut_exact_;
    const double split_margin = node_margin_;  // (the recursion below overwrites the member)
    std::vector<Range> lbox(box);
    lbox[feat].hi = cut;
    lbox[feat].hi_exact = cut_exact;
    int c1 = subdivide(first, first + nleft, lbox, level + 1);
    std::vector<Range> rbox(box);
    rbox[feat].lo = cut;
    rbox[feat].lo_exact = cut_exact;
    int c2 = subdivide(first + nleft, last, rbox, level + 1);
    // census only: are the two plane coordinates numbers both codebooks share bit for bit?  divlow is the largest
    // coordinate of the left child, divhigh the smallest of the right one (span_of: attained only by bit-reproducible
    // points, no other point within rounding noise of it).  Kept in bits 16 / 17 of the node's `a`.
    int div_flags = split_margin <= kKdRobustMargin ? kKdNodeFragile : 0;
    if (exact_) {
      double mn, mx;
      bool mn_e, mx_e;
      span_of(first, nleft, feat, mn, mx, &mn_e, &mx_e);
      if (mx_e) div_flags |= kKdDivLowExact;
      span_of(first + nleft, last - first - nleft, feat, mn, mx, &mn_e, &mx_e);
      if (mn_e) div_flags |= kKdDivHighExact;
    }
    KdNode &nd = out_.nodes[me];
    nd.child1 = c1;
    nd.child2 = c2;
    nd.a = feat | div_flags;
    nd.b = (int)(first + nleft);  // child1 covers order[first, b), child2 order[b, last)
    nd.divlow = lbox[feat].hi;
    nd.divhigh = rbox[feat].lo;
    for (int d = 0; d < dim_; d++) {
      box[d].lo = std::min(lbox[d].lo, rbox[d].lo);
      box[d].hi = std::max(lbox[d].hi, rbox[d].hi);
    }
    return me;
  }

  const double *pts_;
  size_t n_;
  int dim_;
  size_t leaf_max_;
  KdHostTree &out_;
  const unsigned char *exact_;  // per point (may be null): bit-reproducible with either centroid arithmetic
  struct Spread {
    double value;
    int dim;
    bool exact;
  };
  std::vector<Spread> spreads_;  // census scratch of choose_split: the eligible dimensions of the node
  bool cut_exact_ = false;      // of the plane choose_split just returned
  double node_margin_ = 1.0;    // smallest census margin among the comparisons of the node choose_split just handled
};

}  // namespace

void build_kd_tree(const double *points, size_t K, int dim, int leaf_max, KdHostTree &out, const unsigned char *exact) {
  Builder(points, K, dim, leaf_max, out, exact).run();
}

}  // namespace qb
