// Small CUDA plumbing shared by the kernels files: error handling, stream-ordered
// temporary buffers, launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdexcept>
#include <string>

namespace phy
{

struct CudaError : std::runtime_error {
	explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line)
{
	if (e != cudaSuccess) {
		char buf[512];
		snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
		throw CudaError(buf);
	}
}

#define CUDA_CHECK(x) ::phy::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch of the library goes through KERNEL_CHECK; the count is reported as
// the "launches" statistic
inline uint64_t g_kernel_launches = 0;
#define KERNEL_CHECK()                                                                              \
	do {                                                                                            \
		::phy::g_kernel_launches++;                                                                 \
		::phy::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__);                 \
	} while (0)

constexpr int NUM_SMS_B200 = 148;

inline int div_up(int64_t a, int64_t b)
{
	return (int)((a + b - 1) / b);
}

// Stream-ordered device buffer (cudaMallocAsync); freed on the same stream.
template <typename T> class DevBuf
{
	T *p_ = nullptr;
	size_t n_ = 0;
	cudaStream_t s_ = 0;

  public:
	DevBuf() = default;
	DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : p_(o.p_), n_(o.n_), s_(o.s_) { o.p_ = nullptr; o.n_ = 0; }
	DevBuf &operator=(DevBuf &&o) noexcept
	{
		if (this != &o) {
			release();
			p_ = o.p_;
			n_ = o.n_;
			s_ = o.s_;
			o.p_ = nullptr;
			o.n_ = 0;
		}
		return *this;
	}
	~DevBuf() { release(); }
	void alloc(size_t n, cudaStream_t s)
	{
		release();
		s_ = s;
		n_ = n;
		if (n) CUDA_CHECK(cudaMallocAsync((void **)&p_, n * sizeof(T), s));
	}
	void release()
	{
		if (p_) cudaFreeAsync(p_, s_);
		p_ = nullptr;
		n_ = 0;
	}
	void zero() { if (n_) CUDA_CHECK(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s_)); }
	T *get() const { return p_; }
	size_t size() const { return n_; }
	size_t bytes() const { return n_ * sizeof(T); }
	void swap(DevBuf &o)
	{
		std::swap(p_, o.p_);
		std::swap(n_, o.n_);
		std::swap(s_, o.s_);
	}
};

template <typename T> inline T d2h_scalar(const T *dptr, cudaStream_t s)
{
	T v;
	CUDA_CHECK(cudaMemcpyAsync(&v, dptr, sizeof(T), cudaMemcpyDeviceToHost, s));
	CUDA_CHECK(cudaStreamSynchronize(s));
	return v;
}

} // namespace phy
