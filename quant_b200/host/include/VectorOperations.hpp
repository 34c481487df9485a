// Vector types and element-wise arithmetic of the reference's math layer
// (/root/reference/include/VectorOperations.hpp:10-111), restated without Boost.
//
// The reference's `Vector` is boost::container::small_vector<double, 27>: up to 27 elements live
// inside the object, longer vectors spill to the heap.  Boost is only a container there (no
// arithmetic), so this header ships its own small-buffer container with the same surface the
// reference's code and clients use.  Nothing here is on the GPU path: the B200 library works on the
// raw image bytes; these types exist so that code written against Quantizer.hpp / Compressor.hpp
// compiles and behaves unchanged.
#pragma once
#include <algorithm>
#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>

namespace qbhost {

template <class T, std::size_t Inline>
class SmallVec {
 public:
  typedef T value_type;
  typedef T *iterator;
  typedef const T *const_iterator;

  SmallVec() : ptr_(buf_), len_(0), cap_(Inline) {}
  explicit SmallVec(std::size_t n, const T &v = T()) : SmallVec() { resize(n, v); }
  SmallVec(std::initializer_list<T> il) : SmallVec() { for (const T &v : il) push_back(v); }
  SmallVec(const SmallVec &o) : SmallVec() { assign(o.begin(), o.end()); }
  SmallVec(SmallVec &&o) noexcept : SmallVec() { steal(o); }
  template <class It>
  SmallVec(It first, It last) : SmallVec() { assign(first, last); }
  ~SmallVec() { release(); }

  SmallVec &operator=(const SmallVec &o) {
    if (this != &o) assign(o.begin(), o.end());
    return *this;
  }
  SmallVec &operator=(SmallVec &&o) noexcept {
    if (this != &o) { release(); ptr_ = buf_; len_ = 0; cap_ = Inline; steal(o); }
    return *this;
  }

  std::size_t size() const { return len_; }
  bool empty() const { return len_ == 0; }
  T *data() { return ptr_; }
  const T *data() const { return ptr_; }
  iterator begin() { return ptr_; }
  iterator end() { return ptr_ + len_; }
  const_iterator begin() const { return ptr_; }
  const_iterator end() const { return ptr_ + len_; }
  T &operator[](std::size_t i) { return ptr_[i]; }
  const T &operator[](std::size_t i) const { return ptr_[i]; }
  T &at(std::size_t i) { if (i >= len_) throw std::out_of_range("SmallVec::at"); return ptr_[i]; }
  const T &at(std::size_t i) const { if (i >= len_) throw std::out_of_range("SmallVec::at"); return ptr_[i]; }
  T &back() { return ptr_[len_ - 1]; }
  const T &back() const { return ptr_[len_ - 1]; }

  void clear() { len_ = 0; }
  void reserve(std::size_t n) { if (n > cap_) grow(n); }
  void push_back(const T &v) {
    if (len_ == cap_) { T copy = v; grow(cap_ * 2); ptr_[len_++] = copy; } else ptr_[len_++] = v;
  }
  void resize(std::size_t n, const T &v = T()) {
    reserve(n);
    for (std::size_t i = len_; i < n; i++) ptr_[i] = v;
    len_ = n;
  }
  template <class It>
  void assign(It first, It last) { len_ = 0; for (; first != last; ++first) push_back(*first); }
  template <class It>
  void insert(iterator pos, It first, It last) {  // append-or-middle insert (concat uses end())
    std::vector<T> tail(pos, end());
    len_ = static_cast<std::size_t>(pos - ptr_);
    for (; first != last; ++first) push_back(*first);
    for (const T &v : tail) push_back(v);
  }
  bool operator==(const SmallVec &o) const { return len_ == o.len_ && std::equal(begin(), end(), o.begin()); }
  bool operator!=(const SmallVec &o) const { return !(*this == o); }

 private:
  void grow(std::size_t n) {
    T *p = new T[n];
    std::copy(ptr_, ptr_ + len_, p);
    release();
    ptr_ = p;
    cap_ = n;
  }
  void release() { if (ptr_ != buf_) delete[] ptr_; }
  void steal(SmallVec &o) {
    if (o.ptr_ == o.buf_) {
      std::copy(o.ptr_, o.ptr_ + o.len_, buf_);
      len_ = o.len_;
    } else {
      ptr_ = o.ptr_; len_ = o.len_; cap_ = o.cap_;
      o.ptr_ = o.buf_; o.cap_ = Inline;
    }
    o.len_ = 0;
  }
  T *ptr_;
  std::size_t len_, cap_;
  T buf_[Inline];
};

}  // namespace qbhost

// Vector: what the algorithms work on; VectorType: its element type (the reference computes in FP64).
typedef double VectorType;
typedef qbhost::SmallVec<VectorType, 27> Vector;
typedef qbhost::SmallVec<char, 27> CharVector;

namespace qbhost {
template <class F>
inline Vector zip(const Vector &a, const Vector &b, F f) {
  Vector r(a.size());
  for (std::size_t i = 0; i < a.size(); i++) r[i] = f(a[i], b[i]);
  return r;
}
}  // namespace qbhost

static inline Vector operator+(const Vector &a, const Vector &b) { return qbhost::zip(a, b, [](double x, double y) { return x + y; }); }
static inline Vector operator-(const Vector &a, const Vector &b) { return qbhost::zip(a, b, [](double x, double y) { return x - y; }); }
static inline Vector operator*(const Vector &a, const Vector &b) { return qbhost::zip(a, b, [](double x, double y) { return x * y; }); }
static inline Vector operator/(const Vector &a, const Vector &b) { return qbhost::zip(a, b, [](double x, double y) { return x / y; }); }
static inline Vector &operator+=(Vector &a, const Vector &b) { for (std::size_t i = 0; i < a.size(); i++) a[i] += b[i]; return a; }
static inline Vector &operator-=(Vector &a, const Vector &b) { for (std::size_t i = 0; i < a.size(); i++) a[i] -= b[i]; return a; }
static inline Vector operator*(const Vector &a, VectorType s) { Vector r(a); for (auto &v : r) v *= s; return r; }
static inline Vector operator*(VectorType s, const Vector &a) { return a * s; }
static inline Vector operator/(const Vector &a, VectorType s) { Vector r(a); for (auto &v : r) v /= s; return r; }
static inline Vector &operator*=(Vector &a, VectorType s) { for (auto &v : a) v *= s; return a; }
static inline Vector &operator/=(Vector &a, VectorType s) { for (auto &v : a) v /= s; return a; }

// concat(a, b): a followed by b (used by the split step, src/Quantizer.cpp:134).
template <class T>
static inline std::vector<T> concat(const std::vector<T> &a, const std::vector<T> &b) {
  std::vector<T> r;
  r.reserve(a.size() + b.size());
  r.insert(r.end(), a.begin(), a.end());
  r.insert(r.end(), b.begin(), b.end());
  return r;
}

// norm(v): SQUARED Euclidean length, summed left to right (VectorOperations.hpp:107-111).
static inline VectorType norm(const Vector &v) {
  VectorType s = 0;
  for (VectorType x : v) s += x * x;
  return s;
}
