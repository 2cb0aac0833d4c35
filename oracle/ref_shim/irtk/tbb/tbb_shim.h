// tbb_shim.h -- TEST INFRASTRUCTURE.  Stand-in for the few TBB constructs the reference uses (common++/include/irtkParallel.h and
// the Parallel* functors of irtkReconstructionGPU.cc): task_scheduler_init, blocked_range(2d/3d), parallel_for,
// parallel_reduce (splitting constructor + join), tick_count, mutex, concurrent_queue.  Backed by std::thread: a range is cut into
// contiguous chunks, one per thread; parallel_reduce gives every chunk a split-off copy of the body and joins them in order.
// TBB itself is not in this image; no TBB code is contained here.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstddef>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace tbb {

struct split {};

inline int& shim_threads() { static int n = 0; return n; }     // 0 = hardware concurrency
inline int shim_nthreads()
{
    int n = shim_threads();
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    return n > 0 ? n : 1;
}

class task_scheduler_init {
public:
    static const int automatic = -1;
    static const int deferred = -2;
    task_scheduler_init(int n = automatic) { if (n > 0) shim_threads() = n; }
    void initialize(int n = automatic) { if (n > 0) shim_threads() = n; }
    void terminate() {}
    static int default_num_threads() { return (int)std::thread::hardware_concurrency(); }
};

template <class T>
class blocked_range {
    T b_, e_; size_t g_;
public:
    typedef T const_iterator;
    blocked_range() : b_(), e_(), g_(1) {}
    blocked_range(T b, T e, size_t g = 1) : b_(b), e_(e), g_(g) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    size_t size() const { return (size_t)(e_ - b_); }
    size_t grainsize() const { return g_; }
    bool empty() const { return !(b_ < e_); }
};
template <class R, class C = R>
class blocked_range2d {
    blocked_range<R> r_; blocked_range<C> c_;
public:
    blocked_range2d(R rb, R re, size_t, C cb, C ce, size_t) : r_(rb, re), c_(cb, ce) {}
    blocked_range2d(R rb, R re, C cb, C ce) : r_(rb, re), c_(cb, ce) {}
    blocked_range2d(const blocked_range<R>& r, const blocked_range<C>& c) : r_(r), c_(c) {}
    const blocked_range<R>& rows() const { return r_; }
    const blocked_range<C>& cols() const { return c_; }
};
template <class P, class R = P, class C = R>
class blocked_range3d {
    blocked_range<P> p_; blocked_range<R> r_; blocked_range<C> c_;
public:
    blocked_range3d(P pb, P pe, size_t, R rb, R re, size_t, C cb, C ce, size_t) : p_(pb, pe), r_(rb, re), c_(cb, ce) {}
    blocked_range3d(P pb, P pe, R rb, R re, C cb, C ce) : p_(pb, pe), r_(rb, re), c_(cb, ce) {}
    blocked_range3d(const blocked_range<P>& p, const blocked_range<R>& r, const blocked_range<C>& c) : p_(p), r_(r), c_(c) {}
    const blocked_range<P>& pages() const { return p_; }
    const blocked_range<R>& rows() const { return r_; }
    const blocked_range<C>& cols() const { return c_; }
};

namespace shim_detail {
template <class T>
std::vector<blocked_range<T>> chunks(const blocked_range<T>& r)
{
    std::vector<blocked_range<T>> out;
    const size_t n = r.empty() ? 0 : r.size();
    if (n == 0) return out;
    size_t parts = std::min<size_t>((size_t)shim_nthreads() * 4, n);       // a few chunks per thread (dynamic pick-up below)
    if (parts == 0) parts = 1;
    for (size_t p = 0; p < parts; ++p) {
        const T b = r.begin() + (T)(n * p / parts), e = r.begin() + (T)(n * (p + 1) / parts);
        if (b < e) out.push_back(blocked_range<T>(b, e, r.grainsize()));
    }
    return out;
}
template <class F>
void run_indexed(size_t n, F f)
{
    const size_t nt = std::min<size_t>((size_t)shim_nthreads(), n);
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::mutex m; size_t next = 0;
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&] {
            for (;;) {
                size_t i;
                { std::lock_guard<std::mutex> l(m); if (next >= n) return; i = next++; }
                f(i);
            }
        });
    for (auto& x : th) x.join();
}
}  // namespace shim_detail

template <class T, class Body>
void parallel_for(const blocked_range<T>& r, const Body& body)
{
    auto cs = shim_detail::chunks(r);
    shim_detail::run_indexed(cs.size(), [&](size_t i) { body(cs[i]); });
}
template <class R, class C, class Body>
void parallel_for(const blocked_range2d<R, C>& r, const Body& body)
{
    auto cs = shim_detail::chunks(r.rows());
    shim_detail::run_indexed(cs.size(), [&](size_t i) { body(blocked_range2d<R, C>(cs[i], r.cols())); });
}
template <class P, class R, class C, class Body>
void parallel_for(const blocked_range3d<P, R, C>& r, const Body& body)
{
    auto cs = shim_detail::chunks(r.pages());
    shim_detail::run_indexed(cs.size(), [&](size_t i) { body(blocked_range3d<P, R, C>(cs[i], r.rows(), r.cols())); });
}
// one body copy per THREAD (as TBB does when a range is stolen), each fed several chunks, joined in thread order
template <class T, class Body>
void parallel_reduce(const blocked_range<T>& r, Body& body)
{
    auto cs = shim_detail::chunks(r);
    const size_t nt = std::min<size_t>((size_t)shim_nthreads(), cs.size());
    if (nt <= 1) { for (auto& c : cs) body(c); return; }
    std::vector<Body*> bodies(nt, nullptr);
    bodies[0] = &body;
    for (size_t t = 1; t < nt; ++t) bodies[t] = new Body(body, split());
    std::mutex m; size_t next = 0;
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (;;) {
                size_t i;
                { std::lock_guard<std::mutex> l(m); if (next >= cs.size()) return; i = next++; }
                (*bodies[t])(cs[i]);
            }
        });
    for (auto& x : th) x.join();
    for (size_t t = 1; t < nt; ++t) { body.join(*bodies[t]); delete bodies[t]; }
}

class tick_count {
    std::chrono::steady_clock::time_point t_;
public:
    class interval_t {
        double s_;
    public:
        explicit interval_t(double s = 0) : s_(s) {}
        double seconds() const { return s_; }
    };
    static tick_count now() { tick_count t; t.t_ = std::chrono::steady_clock::now(); return t; }
    friend interval_t operator-(const tick_count& a, const tick_count& b) { return interval_t(std::chrono::duration<double>(a.t_ - b.t_).count()); }
};

class mutex {
    std::mutex m_;
public:
    class scoped_lock {
        std::mutex* m_;
    public:
        scoped_lock() : m_(nullptr) {}
        explicit scoped_lock(mutex& m) : m_(&m.m_) { m_->lock(); }
        ~scoped_lock() { if (m_) m_->unlock(); }
        void acquire(mutex& m) { m_ = &m.m_; m_->lock(); }
        void release() { if (m_) { m_->unlock(); m_ = nullptr; } }
    };
    void lock() { m_.lock(); }
    void unlock() { m_.unlock(); }
};

template <class T>
class concurrent_queue {
    std::queue<T> q_; std::mutex m_;
public:
    void push(const T& v) { std::lock_guard<std::mutex> l(m_); q_.push(v); }
    bool try_pop(T& v) { std::lock_guard<std::mutex> l(m_); if (q_.empty()) return false; v = q_.front(); q_.pop(); return true; }
    bool empty() { std::lock_guard<std::mutex> l(m_); return q_.empty(); }
    size_t size() { std::lock_guard<std::mutex> l(m_); return q_.size(); }
    void pop(T& v) { std::lock_guard<std::mutex> l(m_); v = q_.front(); q_.pop(); }
};

template <class T>
class concurrent_bounded_queue : public concurrent_queue<T> {
public:
    void set_capacity(std::ptrdiff_t) {}
};

}  // namespace tbb
