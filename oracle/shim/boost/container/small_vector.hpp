// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for boost::container::small_vector
// so that the *unmodified* reference sources under /root/reference compile in an image that
// ships no Boost headers.  The reference uses the container purely for storage
// (/root/reference/include/VectorOperations.hpp:6,11-12, include/Debug.hpp:4); no arithmetic
// lives in Boost, so this shim cannot change any numeric result.
//
// Storage model follows Boost's: N elements inline, heap beyond that.
#pragma once
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <new>
#include <stdexcept>
#include <type_traits>
#include <utility>

namespace boost {
namespace container {

template <class T, std::size_t N>
class small_vector {
  static_assert(std::is_trivially_copyable<T>::value, "shim supports trivially copyable T only");

 public:
  typedef T value_type;
  typedef T *iterator;
  typedef const T *const_iterator;
  typedef T &reference;
  typedef const T &const_reference;
  typedef std::size_t size_type;
  typedef std::ptrdiff_t difference_type;

  small_vector() : ptr_(inline_ptr()), size_(0), cap_(N) {}
  explicit small_vector(size_type n) : small_vector() { resize(n); }
  small_vector(size_type n, const T &v) : small_vector() {
    reserve(n);
    for (size_type i = 0; i < n; i++) ptr_[i] = v;
    size_ = n;
  }
  small_vector(std::initializer_list<T> il) : small_vector() {
    reserve(il.size());
    for (const T &v : il) ptr_[size_++] = v;
  }
  small_vector(const small_vector &o) : small_vector() { assign_from(o); }
  small_vector(small_vector &&o) noexcept : small_vector() { steal(o); }
  ~small_vector() { release(); }

  small_vector &operator=(const small_vector &o) {
    if (this != &o) {
      size_ = 0;
      assign_from(o);
    }
    return *this;
  }
  small_vector &operator=(small_vector &&o) noexcept {
    if (this != &o) {
      release();
      ptr_ = inline_ptr();
      cap_ = N;
      size_ = 0;
      steal(o);
    }
    return *this;
  }

  size_type size() const { return size_; }
  bool empty() const { return size_ == 0; }
  size_type capacity() const { return cap_; }
  T *data() { return ptr_; }
  const T *data() const { return ptr_; }
  iterator begin() { return ptr_; }
  iterator end() { return ptr_ + size_; }
  const_iterator begin() const { return ptr_; }
  const_iterator end() const { return ptr_ + size_; }
  const_iterator cbegin() const { return ptr_; }
  const_iterator cend() const { return ptr_ + size_; }
  reference operator[](size_type i) { return ptr_[i]; }
  const_reference operator[](size_type i) const { return ptr_[i]; }
  reference at(size_type i) {
    if (i >= size_) throw std::out_of_range("small_vector::at");
    return ptr_[i];
  }
  const_reference at(size_type i) const {
    if (i >= size_) throw std::out_of_range("small_vector::at");
    return ptr_[i];
  }
  reference front() { return ptr_[0]; }
  reference back() { return ptr_[size_ - 1]; }
  const_reference front() const { return ptr_[0]; }
  const_reference back() const { return ptr_[size_ - 1]; }

  void reserve(size_type n) {
    if (n <= cap_) return;
    size_type nc = cap_ * 2 > n ? cap_ * 2 : n;
    T *np = static_cast<T *>(::operator new(nc * sizeof(T)));
    if (size_) std::memcpy(np, ptr_, size_ * sizeof(T));
    release();
    ptr_ = np;
    cap_ = nc;
  }
  void resize(size_type n) {
    reserve(n);
    for (size_type i = size_; i < n; i++) ptr_[i] = T();
    size_ = n;
  }
  void clear() { size_ = 0; }
  void push_back(const T &v) {
    T tmp = v;
    if (size_ == cap_) reserve(size_ + 1);
    ptr_[size_++] = tmp;
  }
  template <class... A>
  void emplace_back(A &&...a) {
    push_back(T(std::forward<A>(a)...));
  }

  friend bool operator==(const small_vector &a, const small_vector &b) {
    if (a.size_ != b.size_) return false;
    for (size_type i = 0; i < a.size_; i++)
      if (!(a.ptr_[i] == b.ptr_[i])) return false;
    return true;
  }
  friend bool operator!=(const small_vector &a, const small_vector &b) { return !(a == b); }

 private:
  T *inline_ptr() { return reinterpret_cast<T *>(inline_); }
  bool is_inline() const { return ptr_ == reinterpret_cast<const T *>(inline_); }
  void release() {
    if (!is_inline()) ::operator delete(ptr_);
  }
  void assign_from(const small_vector &o) {
    reserve(o.size_);
    if (o.size_) std::memcpy(ptr_, o.ptr_, o.size_ * sizeof(T));
    size_ = o.size_;
  }
  void steal(small_vector &o) {
    if (o.is_inline()) {
      if (o.size_) std::memcpy(ptr_, o.ptr_, o.size_ * sizeof(T));
      size_ = o.size_;
    } else {
      ptr_ = o.ptr_;
      cap_ = o.cap_;
      size_ = o.size_;
      o.ptr_ = o.inline_ptr();
      o.cap_ = N;
    }
    o.size_ = 0;
  }

  T *ptr_;
  size_type size_;
  size_type cap_;
  alignas(T) unsigned char inline_[N * sizeof(T)];
};

}  // namespace container
}  // namespace boost
