// psc_api.cu -- the C ABI (include/psc.h): device memory, kernel launches and the host-side
// batch scheduler that replaces the pthread driver of the reference (classes.cpp:285-477).
#include "../../include/psc.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#ifndef PSC_CPU_SHIM
#include <cuda.h> // types only: the driver entry points are looked up at run time (cudaGetDriverEntryPoint)
#endif

#include "psc_internal.h"
#include "k_constraints.cuh"
#include "k_solve.cuh"

using namespace pscdev;

#ifndef PSC_CPU_SHIM
#define PSC_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

// ------------------------------------------------------------------------------------------
// kernels (thin wrappers; the work is in k_constraints.cuh / k_solve.cuh)
// ------------------------------------------------------------------------------------------
struct CoopLaunch {
	int width; // 1 or 32
};

PSC_DEV Coop make_coop(int width)
{
	Coop co;
	co.n = width;
	co.lane = width == 1 ? 0 : (int)(threadIdx.x & 31);
	return co;
}
// index of this cooperative group among all groups of the grid, and their number
PSC_DEV long long coop_id(int width) { return ((long long)blockIdx.x * blockDim.x + threadIdx.x) / width; }
PSC_DEV long long coop_count(int width) { return ((long long)gridDim.x * blockDim.x) / width; }

__global__ void k_group_ranges(GraphDev g, SampleDev s)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, T = (long long)gridDim.x * blockDim.x;
	for (long long c = t; c < g.n_comp; c += T) group_range_one(g, s, (int)c);
}

// units are (sample, component) pairs handed out through a global counter, largest groups first
__global__ void __launch_bounds__(128) k_build_constraints(GraphDev g, const SampleDev *samples, int nSamples,
                                                          const int *order, int nUnits, int *counter,
                                                          int primaryAsUnique, int width)
{
	Coop co = make_coop(width);
	// hash tables / sort lists of a group: a slice of shared memory per warp; the serial CPU harness uses a plain array
	int *smem = nullptr;
#ifndef PSC_CPU_SHIM
	__shared__ int cg_smem[4 * PSC_CG_SMEM_INTS];
	if (width == 32) smem = cg_smem + (threadIdx.x >> 5) * PSC_CG_SMEM_INTS;
#else
	static thread_local int cg_host[PSC_CG_SMEM_INTS];
	smem = cg_host;
#endif
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nUnits) break;
		int unit = order ? order[u] : u;
		int smp = unit / g.n_comp, c = unit % g.n_comp;
		build_group(co, g, samples[smp], c, primaryAsUnique != 0, smem);
		co.sync();
	}
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
	bool ensure(size_t n)
	{
		if (n <= bytes && p) return true;
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
		if (cudaMalloc(&p, n ? n : 16) != cudaSuccess) return false;
		bytes = n ? n : 16;
		return true;
	}
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		bytes = 0;
	}
	template <class T> T *as() { return (T *)p; }
};

// pinned host staging memory (device -> host copies of results run at full PCIe speed only from pinned memory)
struct HostBuf {
	void *p = nullptr;
	size_t bytes = 0;
	bool ensure(size_t n)
	{
		if (n <= bytes && p) return true;
		if (p) cudaFreeHost(p);
		p = nullptr;
		bytes = 0;
		size_t want = n + n / 4 + 4096;
		if (cudaMallocHost(&p, want) != cudaSuccess) return false;
		bytes = want;
		return true;
	}
	void release()
	{
		if (p) cudaFreeHost(p);
		p = nullptr;
		bytes = 0;
	}
};

#include "psc_solve_state.h"

__global__ void k_pow_quarter(const double *x, double *y, long long n)
{
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#ifdef PSC_CPU_SHIM
	for (i = 0; i < n; ++i) y[i] = pscdev::pow_quarter(x[i]);
#else
	if (i < n) y[i] = pscdev::pow_quarter(x[i]);
#endif
}

struct SampleState {
	SampleDev d;
	bool loaded = false;
	DevBuf chr_off, pos, mate_pos, flags, nh, name_key, seg_off, seg_a, seg_b;
	DevBuf grp_lo, grp_hi, bits_off, n_con, n_pair;
	DevBuf cand_meta, cand_cid, cand_bits, sortbuf, pair_key, pair_uniq_in;
	DevBuf con_bits, con_first, con_last, con_support, con_uniq, con_maxlen, con_nhcnt, con_weight, con_norm;
	DevBuf pair_i, pair_j, pair_support, pair_uniq, pair_ec, pair_norm;
	std::vector<int> h_lo, h_hi, h_ncon, h_npair;
	std::vector<long long> h_bits_off;
	void release()
	{
		DevBuf *all[] = {&chr_off, &pos, &mate_pos, &flags, &nh, &name_key, &seg_off, &seg_a, &seg_b, &grp_lo,
		                 &grp_hi, &bits_off, &n_con, &n_pair, &cand_meta, &cand_cid, &cand_bits, &sortbuf,
		                 &pair_key, &pair_uniq_in, &con_bits, &con_first, &con_last, &con_support, &con_uniq,
		                 &con_maxlen, &con_nhcnt, &con_weight, &con_norm, &pair_i, &pair_j, &pair_support,
		                 &pair_uniq, &pair_ec, &pair_norm};
		for (DevBuf *b : all) b->release();
		loaded = false;
	}
};

struct psc_ctx {
	int device = 0;
	cudaStream_t stream = 0;
	cudaStream_t stream2 = 0; // per-unit kernels of the enumerated components, running under the DP kernel's tail
	// psc_solve's streams.  When the driver offers green contexts the SMs are split into a small partition for the
	// few critical (component, sample) units whose single-warp latency chains set the step time (they then run one
	// warp per SM sub-partition, undisturbed) and the rest for everything else: sMain / sSecond live in the big
	// partition, sCrit in the small one.  Otherwise these are ordinary streams (sMain == stream, sSecond == stream2).
	cudaStream_t sMain = 0, sSecond = 0, sCrit = 0, sThird = 0; // sThird: the chain of the enumerated components
	cudaEvent_t evCritGo = 0, evCritDone = 0, evT1Go = 0, evT1Done = 0;
	int critSms = 0; // SMs of the critical partition, 0: no partitioning
#ifndef PSC_CPU_SHIM
	CUgreenCtx gctxCrit = nullptr, gctxRest = nullptr;
#endif
	cudaEvent_t ev0 = 0, ev1 = 0;
	std::string err;
	psc_params prm;
	bool have_params = false, have_graph = false, constraints_built = false, solved = false;
	int coop = 32;
	int n_sm = 148;
	size_t mem_budget = (size_t)32 << 30;

	// graph: host copies needed for sizing + device arrays
	GraphDev gd;
	std::vector<int> h_comp_off, h_comp_chr;
	DevBuf g_comp_off, g_comp_chr, g_comp_base, g_comp_used, g_comp_max_end, g_comp_prev_end, g_se_start,
	    g_se_end, g_se_pmax, g_lt, g_rt, g_ls, g_rs, g_lcls, g_rcls, g_gene, g_next_off, g_next_idx, g_prev_off,
	    g_prev_idx;
	std::vector<int> h_next_cnt; // per component: number of next edges
	std::vector<int> h_comp_ng;  // per component: usedGeneId - baseGeneId

	std::vector<SampleState> samples;
	DevBuf d_samples, d_order, d_counter;

	SolveState solve;

	psc_timing timing;
	// per-launch CUDA events (resolved at the next synchronisation point)
	std::vector<cudaEvent_t> evPool;
	std::vector<int> evIds;
	size_t evUsed = 0;
	cudaStream_t evStream = 0;
	void ktime_begin(int id, cudaStream_t on = nullptr)
	{
		evStream = on ? on : stream;
		if (evUsed + 2 > evPool.size()) {
			evPool.resize(evUsed + 2);
			cudaEventCreate(&evPool[evUsed]);
			cudaEventCreate(&evPool[evUsed + 1]);
		}
		evIds.push_back(id);
		cudaEventRecord(evPool[evUsed], evStream);
	}
	void ktime_end()
	{
		cudaEventRecord(evPool[evUsed + 1], evStream);
		evUsed += 2;
	}
	void ktime_collect() // call after a stream synchronisation
	{
		for (size_t k = 0; k < evIds.size(); ++k) {
			float ms = 0;
			cudaEventElapsedTime(&ms, evPool[2 * k], evPool[2 * k + 1]);
			timing.kernel_ms[evIds[k]] += ms;
			timing.kernel_count[evIds[k]] += 1;
		}
		evIds.clear();
		evUsed = 0;
	}
	// psc_solve drives its two streams from two host threads: launches (with their timing events) and the error
	// string are serialised
	std::mutex launchMu, errMu;
	int fail(const std::string &m)
	{
		std::lock_guard<std::mutex> lk(errMu);
		err = m;
		return 1;
	}
};

#define KLAUNCH(id, kernel, grid, block, ...)                      \
	do {                                                            \
		ctx->ktime_begin(id);                                        \
		PSC_LAUNCH(kernel, grid, block, ctx->stream, __VA_ARGS__);   \
		ctx->ktime_end();                                            \
	} while (0)

#define KLAUNCH_ON(strm, id, kernel, grid, block, ...)             \
	do {                                                            \
		std::lock_guard<std::mutex> lk_(ctx->launchMu);              \
		ctx->ktime_begin(id, strm);                                  \
		PSC_LAUNCH(kernel, grid, block, strm, __VA_ARGS__);          \
		ctx->ktime_end();                                            \
	} while (0)

#define CK(call)                                                                                   \
	do {                                                                                            \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess) return ctx->fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

template <class T> static bool upload(psc_ctx *ctx, DevBuf &b, const T *src, size_t n, cudaStream_t on = nullptr)
{
	if (!b.ensure(n * sizeof(T))) return false;
	if (n) return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, on ? on : ctx->stream) == cudaSuccess;
	return true;
}

extern "C" {

int psc_create(psc_ctx **out, int device)
{
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return 2; // no CUDA device: there is no CPU fallback
	if (device < 0 || device >= n) return 3;
	if (cudaSetDevice(device) != cudaSuccess) return 4;
	cudaDeviceSetLimit(cudaLimitStackSize, 4096); // the DP state machine keeps a private copy of its state on the stack
	psc_ctx *ctx = new psc_ctx;
	ctx->device = device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount;
	size_t fr = 0, tot = 0;
	if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->mem_budget = fr / 4;
	// scratch budget per buffer class; work that needs more is cut into chunks of components (tests force tiny budgets)
	if (const char *e = getenv("PSC_MEM_BUDGET_MB")) ctx->mem_budget = (size_t)atoll(e) << 20;
	if (cudaStreamCreate(&ctx->stream) != cudaSuccess || cudaStreamCreate(&ctx->stream2) != cudaSuccess ||
	    cudaEventCreate(&ctx->ev0) != cudaSuccess ||
	    cudaEventCreate(&ctx->ev1) != cudaSuccess) {
		delete ctx;
		return 5;
	}
	memset(&ctx->timing, 0, sizeof(ctx->timing));
	memset(&ctx->prm, 0, sizeof(ctx->prm));
	cudaEventCreateWithFlags(&ctx->evCritGo, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->evCritDone, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->evT1Go, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->evT1Done, cudaEventDisableTiming);
	ctx->sMain = ctx->stream;
	ctx->sSecond = ctx->stream2;
#ifndef PSC_CPU_SHIM
	// SM partitioning (PSC_CRIT_SMS=<n> SMs for the critical partition, default 16; 0 disables)
	{
		int want = 16;
		if (const char *e = getenv("PSC_CRIT_SMS")) want = atoi(e);
		if (want > 0 && ctx->n_sm >= 4 * want) {
			typedef CUresult (*fnDevGet)(CUdevice *, int);
			typedef CUresult (*fnGetRes)(CUdevice, CUdevResource *, CUdevResourceType);
			typedef CUresult (*fnSplit)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
			typedef CUresult (*fnDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
			typedef CUresult (*fnCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
			typedef CUresult (*fnStream)(CUstream *, CUgreenCtx, unsigned int, int);
			auto sym = [](const char *name) -> void * {
				void *f = nullptr;
				cudaDriverEntryPointQueryResult q;
				if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
				return f;
			};
			fnDevGet devGet = (fnDevGet)sym("cuDeviceGet");
			fnGetRes getRes = (fnGetRes)sym("cuDeviceGetDevResource");
			fnSplit split = (fnSplit)sym("cuDevSmResourceSplitByCount");
			fnDesc mkDesc = (fnDesc)sym("cuDevResourceGenerateDesc");
			fnCreate mkCtx = (fnCreate)sym("cuGreenCtxCreate");
			fnStream mkStream = (fnStream)sym("cuGreenCtxStreamCreate");
			bool ok = devGet && getRes && split && mkDesc && mkCtx && mkStream;
			CUdevice dev = 0;
			CUdevResource all, crit, rest;
			unsigned int groups = 1;
			CUdevResourceDesc dCrit = nullptr, dRest = nullptr;
			CUstream a = nullptr, b = nullptr, c = nullptr, d = nullptr;
			ok = ok && devGet(&dev, device) == CUDA_SUCCESS && getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
			     split(&crit, &groups, &all, &rest, 0, (unsigned int)want) == CUDA_SUCCESS && groups == 1 && rest.sm.smCount > 0 &&
			     mkDesc(&dCrit, &crit, 1) == CUDA_SUCCESS && mkDesc(&dRest, &rest, 1) == CUDA_SUCCESS &&
			     mkCtx(&ctx->gctxCrit, dCrit, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
			     mkCtx(&ctx->gctxRest, dRest, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
			     mkStream(&a, ctx->gctxRest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
			     mkStream(&b, ctx->gctxRest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
			     mkStream(&c, ctx->gctxCrit, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
			     mkStream(&d, ctx->gctxRest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
			if (ok) {
				ctx->sMain = (cudaStream_t)a;
				ctx->sSecond = (cudaStream_t)b;
				ctx->sCrit = (cudaStream_t)c;
				ctx->sThird = (cudaStream_t)d;
				ctx->critSms = (int)crit.sm.smCount;
				if (getenv("PSC_DEBUG_HOST"))
					fprintf(stderr, "[host] SM partitions: %u critical + %u\n", crit.sm.smCount, rest.sm.smCount);
			} else
				cudaGetLastError();
		}
	}
#endif
	if ((!ctx->sCrit && cudaStreamCreate(&ctx->sCrit) != cudaSuccess) || (!ctx->sThird && cudaStreamCreate(&ctx->sThird) != cudaSuccess)) {
		delete ctx;
		return 5;
	}
	*out = ctx;
	return 0;
}

void psc_destroy(psc_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (auto &s : ctx->samples) s.release();
	DevBuf *all[] = {&ctx->g_comp_off, &ctx->g_comp_chr, &ctx->g_comp_base, &ctx->g_comp_used,
	                 &ctx->g_comp_max_end, &ctx->g_comp_prev_end, &ctx->g_se_start, &ctx->g_se_end, &ctx->g_se_pmax,
	                 &ctx->g_lt, &ctx->g_rt, &ctx->g_ls, &ctx->g_rs, &ctx->g_lcls, &ctx->g_rcls, &ctx->g_gene,
	                 &ctx->g_next_off, &ctx->g_next_idx, &ctx->g_prev_off, &ctx->g_prev_idx, &ctx->d_samples,
	                 &ctx->d_order, &ctx->d_counter};
	for (DevBuf *b : all) b->release();
	ctx->solve.release();
	for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
	cudaEventDestroy(ctx->ev0);
	cudaEventDestroy(ctx->ev1);
	cudaEventDestroy(ctx->evCritGo);
	cudaEventDestroy(ctx->evCritDone);
	cudaStreamDestroy(ctx->sCrit);
	cudaStreamDestroy(ctx->sThird);
	cudaEventDestroy(ctx->evT1Go);
	cudaEventDestroy(ctx->evT1Done);
	if (ctx->sMain != ctx->stream) cudaStreamDestroy(ctx->sMain);
	if (ctx->sSecond != ctx->stream2) cudaStreamDestroy(ctx->sSecond);
#ifndef PSC_CPU_SHIM
	{
		typedef CUresult (*fnDestroy)(CUgreenCtx);
		void *f = nullptr;
		cudaDriverEntryPointQueryResult q;
		if ((ctx->gctxCrit || ctx->gctxRest) && cudaGetDriverEntryPoint("cuGreenCtxDestroy", &f, cudaEnableDefault, &q) == cudaSuccess && f) {
			if (ctx->gctxCrit) ((fnDestroy)f)(ctx->gctxCrit);
			if (ctx->gctxRest) ((fnDestroy)f)(ctx->gctxRest);
		}
	}
#endif
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->stream2);
	delete ctx;
}

const char *psc_last_error(const psc_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int psc_set_coop_width(psc_ctx *ctx, int32_t width)
{
	if (width != 1 && width != 32) return ctx->fail("coop width must be 1 or 32");
#ifdef PSC_CPU_SHIM
	width = 1;
#endif
	ctx->coop = width;
	return 0;
}

int psc_debug_pow_quarter(psc_ctx *ctx, const double *x, double *y, int64_t n)
{
	cudaSetDevice(ctx->device);
	if (n <= 0) return 0;
	DevBuf dx, dy;
	if (!dx.ensure(sizeof(double) * (size_t)n) || !dy.ensure(sizeof(double) * (size_t)n)) return ctx->fail("out of device memory");
	CK(cudaMemcpyAsync(dx.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
	PSC_LAUNCH(k_pow_quarter, (int)((n + 255) / 256), 256, ctx->stream, dx.as<double>(), dy.as<double>(), (long long)n);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(y, dy.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	dx.release();
	dy.release();
	return 0;
}

int psc_set_params(psc_ctx *ctx, const psc_params *p)
{
	if (p->n_samples <= 0) return ctx->fail("n_samples must be positive");
	ctx->prm = *p;
	ctx->have_params = true;
	for (auto &s : ctx->samples) s.release();
	ctx->samples.clear();
	ctx->samples.resize((size_t)p->n_samples);
	ctx->constraints_built = ctx->solved = false;
#ifdef PSC_CPU_SHIM
	ctx->coop = 1;
#endif
	return 0;
}

int psc_set_graph(psc_ctx *ctx, const psc_graph *g)
{
	cudaSetDevice(ctx->device);
	int nc = g->n_comp, ns = g->n_se;
	if (nc < 0 || ns < 0) return ctx->fail("bad graph sizes");
	ctx->h_comp_off.assign(g->comp_off, g->comp_off + nc + 1);
	ctx->h_comp_chr.assign(g->comp_chr, g->comp_chr + nc);
	std::vector<int> maxEnd((size_t)nc), prevEnd((size_t)nc), pmax((size_t)ns);
	ctx->h_next_cnt.assign((size_t)nc, 0);
	ctx->h_comp_ng.assign((size_t)nc, 0);
	for (int c = 0; c < nc; ++c) {
		int m = INT_MIN;
		for (int i = g->comp_off[c]; i < g->comp_off[c + 1]; ++i) {
			m = std::max(m, g->se_end[i]);
			pmax[(size_t)i] = m;
		}
		maxEnd[(size_t)c] = m;
		prevEnd[(size_t)c] = (c > 0 && g->comp_chr[c - 1] == g->comp_chr[c]) ? maxEnd[(size_t)c - 1] : INT_MIN;
		if (c > 0 && g->comp_chr[c - 1] > g->comp_chr[c]) return ctx->fail("components are not in genome order");
		ctx->h_next_cnt[(size_t)c] = g->next_off[g->comp_off[c + 1]] - g->next_off[g->comp_off[c]];
		ctx->h_comp_ng[(size_t)c] = std::max(0, g->comp_used_gene[c] - g->comp_base_gene[c]);
	}
	bool ok = upload(ctx, ctx->g_comp_off, g->comp_off, (size_t)nc + 1) &&
	          upload(ctx, ctx->g_comp_chr, g->comp_chr, (size_t)nc) &&
	          upload(ctx, ctx->g_comp_base, g->comp_base_gene, (size_t)nc) &&
	          upload(ctx, ctx->g_comp_used, g->comp_used_gene, (size_t)nc) &&
	          upload(ctx, ctx->g_comp_max_end, maxEnd.data(), (size_t)nc) &&
	          upload(ctx, ctx->g_comp_prev_end, prevEnd.data(), (size_t)nc) &&
	          upload(ctx, ctx->g_se_start, g->se_start, (size_t)ns) && upload(ctx, ctx->g_se_end, g->se_end, (size_t)ns) &&
	          upload(ctx, ctx->g_se_pmax, pmax.data(), (size_t)ns) && upload(ctx, ctx->g_lt, g->se_left_type, (size_t)ns) &&
	          upload(ctx, ctx->g_rt, g->se_right_type, (size_t)ns) && upload(ctx, ctx->g_ls, g->se_left_strand, (size_t)ns) &&
	          upload(ctx, ctx->g_rs, g->se_right_strand, (size_t)ns) && upload(ctx, ctx->g_lcls, g->se_left_cls, (size_t)ns) &&
	          upload(ctx, ctx->g_rcls, g->se_right_cls, (size_t)ns) && upload(ctx, ctx->g_gene, g->se_gene_id, (size_t)ns) &&
	          upload(ctx, ctx->g_next_off, g->next_off, (size_t)ns + 1) &&
	          upload(ctx, ctx->g_next_idx, g->next_idx, (size_t)g->next_off[ns]) &&
	          upload(ctx, ctx->g_prev_off, g->prev_off, (size_t)ns + 1) &&
	          upload(ctx, ctx->g_prev_idx, g->prev_idx, (size_t)g->prev_off[ns]);
	if (!ok) return ctx->fail("graph upload failed (out of device memory?)");
	CK(cudaStreamSynchronize(ctx->stream)); // maxEnd/prevEnd/pmax are stack-local staging buffers
	GraphDev &d = ctx->gd;
	d.n_comp = nc;
	d.n_se = ns;
	d.comp_off = ctx->g_comp_off.as<int>();
	d.comp_chr = ctx->g_comp_chr.as<int>();
	d.comp_base = ctx->g_comp_base.as<int>();
	d.comp_used = ctx->g_comp_used.as<int>();
	d.comp_max_end = ctx->g_comp_max_end.as<int>();
	d.comp_prev_end = ctx->g_comp_prev_end.as<int>();
	d.se_start = ctx->g_se_start.as<int>();
	d.se_end = ctx->g_se_end.as<int>();
	d.se_pmax_end = ctx->g_se_pmax.as<int>();
	d.lt = ctx->g_lt.as<int8_t>();
	d.rt = ctx->g_rt.as<int8_t>();
	d.ls = ctx->g_ls.as<int8_t>();
	d.rs = ctx->g_rs.as<int8_t>();
	d.lcls = ctx->g_lcls.as<double>();
	d.rcls = ctx->g_rcls.as<double>();
	d.gene = ctx->g_gene.as<int>();
	d.next_off = ctx->g_next_off.as<int>();
	d.next_idx = ctx->g_next_idx.as<int>();
	d.prev_off = ctx->g_prev_off.as<int>();
	d.prev_idx = ctx->g_prev_idx.as<int>();
	ctx->have_graph = true;
	ctx->constraints_built = ctx->solved = false;
	return 0;
}

int psc_set_reads(psc_ctx *ctx, int32_t sample, const psc_reads *r)
{
	cudaSetDevice(ctx->device);
	if (!ctx->have_params) return ctx->fail("psc_set_params must come first");
	if (sample < 0 || sample >= ctx->prm.n_samples) return ctx->fail("sample index out of range");
	if (r->n_reads >= ((long long)1 << 30)) return ctx->fail("more than 2^30 reads in one sample");
	SampleState &s = ctx->samples[(size_t)sample];
	size_t n = (size_t)r->n_reads;
	bool ok = upload(ctx, s.chr_off, r->chr_off, (size_t)r->n_chr + 1) && upload(ctx, s.pos, r->pos, n) &&
	          upload(ctx, s.mate_pos, r->mate_pos, n) && upload(ctx, s.flags, r->flags, n) &&
	          upload(ctx, s.nh, r->nh, n) && upload(ctx, s.name_key, r->name_key, n) &&
	          upload(ctx, s.seg_off, r->seg_off, n + 1) && upload(ctx, s.seg_a, r->seg_a, (size_t)r->n_segs) &&
	          upload(ctx, s.seg_b, r->seg_b, (size_t)r->n_segs);
	if (!ok) return ctx->fail("reads upload failed (out of device memory?)");
	SampleDev &d = s.d;
	memset(&d, 0, sizeof(d));
	d.n_reads = r->n_reads;
	d.n_segs = r->n_segs;
	d.n_chr = r->n_chr;
	d.chr_off = s.chr_off.as<long long>();
	d.pos = s.pos.as<int>();
	d.mate_pos = s.mate_pos.as<int>();
	d.flags = s.flags.as<uint8_t>();
	d.nh = s.nh.as<uint8_t>();
	d.name_key = s.name_key.as<uint64_t>();
	d.seg_off = s.seg_off.as<uint32_t>();
	d.seg_a = s.seg_a.as<int>();
	d.seg_b = s.seg_b.as<int>();
	d.read_len = r->read_len;
	d.frag_stdev = r->frag_stdev;
	s.loaded = true;
	ctx->constraints_built = ctx->solved = false;
	return 0;
}

int psc_build_constraints(psc_ctx *ctx)
{
	cudaSetDevice(ctx->device);
	if (!ctx->have_graph || !ctx->have_params) return ctx->fail("graph and params must be set first");
	int S = ctx->prm.n_samples, nc = ctx->gd.n_comp;
	for (auto &s : ctx->samples)
		if (!s.loaded) return ctx->fail("psc_set_reads missing for a sample");
	CK(cudaEventRecord(ctx->ev0, ctx->stream));
	long long launches = 0;
	memset(ctx->timing.kernel_ms, 0, sizeof(ctx->timing.kernel_ms));
	memset(ctx->timing.kernel_count, 0, sizeof(ctx->timing.kernel_count));
	memset(ctx->timing.kernel_bytes, 0, sizeof(ctx->timing.kernel_bytes));
	// ---- K0: group ranges ----
	for (int si = 0; si < S; ++si) {
		SampleState &s = ctx->samples[(size_t)si];
		size_t n = (size_t)s.d.n_reads;
		bool ok = s.grp_lo.ensure(sizeof(int) * (size_t)(nc + 1)) && s.grp_hi.ensure(sizeof(int) * (size_t)(nc + 1)) &&
		          s.bits_off.ensure(sizeof(long long) * (size_t)(nc + 1)) && s.n_con.ensure(sizeof(int) * (size_t)(nc + 1)) &&
		          s.n_pair.ensure(sizeof(int) * (size_t)(nc + 1)) && s.cand_meta.ensure(sizeof(int4) * n) &&
		          s.cand_cid.ensure(sizeof(int) * n) && s.sortbuf.ensure(sizeof(int) * (4 * n + 64)) &&
		          s.pair_key.ensure(sizeof(uint64_t) * n) && s.pair_uniq_in.ensure(sizeof(int) * n) &&
		          s.con_first.ensure(sizeof(int) * n) && s.con_last.ensure(sizeof(int) * n) &&
		          s.con_support.ensure(sizeof(int) * n) && s.con_uniq.ensure(sizeof(int) * n) &&
		          s.con_maxlen.ensure(sizeof(int) * n) && s.con_nhcnt.ensure(sizeof(int4) * n) &&
		          s.con_weight.ensure(sizeof(double) * n) && s.con_norm.ensure(sizeof(double) * n) &&
		          s.pair_i.ensure(sizeof(int) * n) && s.pair_j.ensure(sizeof(int) * n) &&
		          s.pair_support.ensure(sizeof(int) * n) && s.pair_uniq.ensure(sizeof(int) * n) &&
		          s.pair_ec.ensure(sizeof(int) * n) && s.pair_norm.ensure(sizeof(double) * n);
		if (!ok) return ctx->fail("out of device memory (constraint tables)");
		SampleDev &d = s.d;
		d.grp_lo = s.grp_lo.as<int>();
		d.grp_hi = s.grp_hi.as<int>();
		d.bits_off = s.bits_off.as<long long>();
		d.n_con = s.n_con.as<int>();
		d.n_pair = s.n_pair.as<int>();
		d.cand_meta = s.cand_meta.as<int4>();
		d.cand_cid = s.cand_cid.as<int>();
		d.sortbuf = s.sortbuf.as<int>();
		d.pair_key = s.pair_key.as<uint64_t>();
		d.pair_uniq_in = s.pair_uniq_in.as<int>();
		d.con_first = s.con_first.as<int>();
		d.con_last = s.con_last.as<int>();
		d.con_support = s.con_support.as<int>();
		d.con_uniq = s.con_uniq.as<int>();
		d.con_maxlen = s.con_maxlen.as<int>();
		d.con_nhcnt = s.con_nhcnt.as<int4>();
		d.con_weight = s.con_weight.as<double>();
		d.con_norm = s.con_norm.as<double>();
		d.pair_i = s.pair_i.as<int>();
		d.pair_j = s.pair_j.as<int>();
		d.pair_support = s.pair_support.as<int>();
		d.pair_uniq = s.pair_uniq.as<int>();
		d.pair_ec = s.pair_ec.as<int>();
		d.pair_norm = s.pair_norm.as<double>();
		if (nc > 0) {
			KLAUNCH(0, k_group_ranges, (nc + 255) / 256, 256, ctx->gd, d);
			++launches;
		}
	}
	CK(cudaGetLastError());
	// ---- offsets of the bit rows (needs the totals on the host to size the buffers) ----
	std::vector<int> order;
	std::vector<long long> cost;
	long long algBytes = 0;
	for (int si = 0; si < S; ++si) {
		SampleState &s = ctx->samples[(size_t)si];
		s.h_lo.resize((size_t)nc);
		s.h_hi.resize((size_t)nc);
		s.h_bits_off.resize((size_t)nc + 1);
		CK(cudaMemcpyAsync(s.h_lo.data(), s.d.grp_lo, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(s.h_hi.data(), s.d.grp_hi, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	for (int si = 0; si < S; ++si) {
		SampleState &s = ctx->samples[(size_t)si];
		long long acc = 0;
		for (int c = 0; c < nc; ++c) {
			int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
			s.h_bits_off[(size_t)c] = acc;
			acc += (long long)(s.h_hi[(size_t)c] - s.h_lo[(size_t)c]) * ((n + 63) / 64);
		}
		s.h_bits_off[(size_t)nc] = acc;
		if (!s.cand_bits.ensure(sizeof(uint64_t) * (size_t)(acc + 1)) || !s.con_bits.ensure(sizeof(uint64_t) * (size_t)(acc + 1)))
			return ctx->fail("out of device memory (bit rows)");
		s.d.cand_bits = s.cand_bits.as<uint64_t>();
		s.d.con_bits = s.con_bits.as<uint64_t>();
		CK(cudaMemcpyAsync(s.d.bits_off, s.h_bits_off.data(), sizeof(long long) * (size_t)(nc + 1), cudaMemcpyHostToDevice,
		                   ctx->stream));
		// algorithmic bytes of the constraint kernel: read records once, write the tables once
		algBytes += s.d.n_reads * 22 + s.d.n_segs * 8;
	}
	// ---- K1+K2: largest groups first (they are the tail of the launch otherwise) ----
	int nUnits = S * nc;
	order.resize((size_t)nUnits);
	for (int u = 0; u < nUnits; ++u) order[(size_t)u] = u;
	{
		std::vector<int> sz((size_t)nUnits);
		for (int si = 0; si < S; ++si)
			for (int c = 0; c < nc; ++c)
				sz[(size_t)si * nc + c] = ctx->samples[(size_t)si].h_hi[(size_t)c] - ctx->samples[(size_t)si].h_lo[(size_t)c];
		// largest first is a scheduling hint: buckets by bit length instead of a comparison sort of S * n_comp units
		size_t cnt[34] = {0};
		std::vector<unsigned char> bucket((size_t)nUnits);
		for (int u = 0; u < nUnits; ++u) {
			int v = sz[(size_t)u], b = 0;
			while (v > 0) { v >>= 1; ++b; }
			bucket[(size_t)u] = (unsigned char)(32 - b);
			++cnt[bucket[(size_t)u] + 1];
		}
		for (int b = 0; b < 33; ++b) cnt[b + 1] += cnt[b];
		for (int u = 0; u < nUnits; ++u) order[cnt[bucket[(size_t)u]]++] = u;
	}
	std::vector<SampleDev> hs((size_t)S);
	for (int si = 0; si < S; ++si) hs[(size_t)si] = ctx->samples[(size_t)si].d;
	if (!upload(ctx, ctx->d_samples, hs.data(), (size_t)S) || !upload(ctx, ctx->d_order, order.data(), (size_t)nUnits) ||
	    !ctx->d_counter.ensure(256))
		return ctx->fail("out of device memory");
	CK(cudaMemsetAsync(ctx->d_counter.p, 0, 256, ctx->stream));
	if (nUnits > 0) {
		int block = 128;
		int grid = ctx->n_sm * 16;
		KLAUNCH(1, k_build_constraints, grid, block, ctx->gd, ctx->d_samples.as<SampleDev>(), S, ctx->d_order.as<int>(), nUnits,
		        ctx->d_counter.as<int>(), ctx->prm.primary_as_unique, ctx->coop);
		++launches;
	}
	CK(cudaGetLastError());
	for (int si = 0; si < S; ++si) {
		SampleState &s = ctx->samples[(size_t)si];
		s.h_ncon.resize((size_t)nc);
		s.h_npair.resize((size_t)nc);
		CK(cudaMemcpyAsync(s.h_ncon.data(), s.d.n_con, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(s.h_npair.data(), s.d.n_pair, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CK(cudaEventRecord(ctx->ev1, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream)); // also keeps hs/order alive until the copies are done
	float ms = 0;
	cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
	ctx->ktime_collect();
	for (int si = 0; si < S; ++si) {
		SampleState &s = ctx->samples[(size_t)si];
		for (int c = 0; c < nc; ++c) {
			int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
			algBytes += (long long)s.h_ncon[(size_t)c] * (8 * ((n + 63) / 64) + 40) + (long long)s.h_npair[(size_t)c] * 24;
		}
	}
	ctx->timing.constraints_ms = ms;
	ctx->timing.kernel_launches = launches;
	ctx->timing.constraint_bytes = algBytes;
	ctx->timing.kernel_bytes[1] = algBytes;
	ctx->constraints_built = true;
	ctx->solved = false;
	// The per-read scratch rows (candidate bit rows, hash-table rows, join keys: ~56 bytes + a bit row per read) are only
	// needed while the tables are built.  They stay allocated for the next call unless memory is short (less than their
	// own size left free) or PSC_FREE_SCRATCH=1: then psc_solve gets the room instead.
	{
		size_t scratch = 0;
		for (auto &s : ctx->samples) scratch += s.cand_meta.bytes + s.cand_cid.bytes + s.cand_bits.bytes + s.sortbuf.bytes + s.pair_key.bytes + s.pair_uniq_in.bytes;
		size_t fr = 0, tot = 0;
		const char *e = getenv("PSC_FREE_SCRATCH");
		bool release = e ? atoi(e) != 0 : (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr < scratch);
		if (release)
			for (auto &s : ctx->samples) {
				s.cand_meta.release(); s.cand_cid.release(); s.cand_bits.release(); s.sortbuf.release(); s.pair_key.release(); s.pair_uniq_in.release();
			}
	}
	return 0;
}

int psc_constraint_counts(psc_ctx *ctx, int32_t *n_constraints, int32_t *n_pairs)
{
	if (!ctx->constraints_built) return ctx->fail("psc_build_constraints has not run");
	int S = ctx->prm.n_samples, nc = ctx->gd.n_comp;
	for (int si = 0; si < S; ++si)
		for (int c = 0; c < nc; ++c) {
			n_constraints[(size_t)si * nc + c] = ctx->samples[(size_t)si].h_ncon[(size_t)c];
			n_pairs[(size_t)si * nc + c] = ctx->samples[(size_t)si].h_npair[(size_t)c];
		}
	return 0;
}

int psc_get_constraint_table(psc_ctx *ctx, int32_t sample, int32_t comp, psc_constraint_table *out)
{
	cudaSetDevice(ctx->device);
	if (!ctx->constraints_built) return ctx->fail("psc_build_constraints has not run");
	if (sample < 0 || sample >= ctx->prm.n_samples || comp < 0 || comp >= ctx->gd.n_comp)
		return ctx->fail("index out of range");
	SampleState &s = ctx->samples[(size_t)sample];
	int C = s.h_ncon[(size_t)comp], M = s.h_npair[(size_t)comp], lo = s.h_lo[(size_t)comp];
	int n = ctx->h_comp_off[(size_t)comp + 1] - ctx->h_comp_off[(size_t)comp];
	int W = (n + 63) / 64;
	out->n_constraints = C;
	out->n_pairs = M;
	out->n_words = W;
	auto dl = [&](void *dst, const void *src, size_t bytes) {
		return !dst || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
	};
	bool ok = dl(out->bits, s.d.con_bits + s.h_bits_off[(size_t)comp], sizeof(uint64_t) * (size_t)C * W) &&
	          dl(out->first, s.d.con_first + lo, sizeof(int) * (size_t)C) && dl(out->last, s.d.con_last + lo, sizeof(int) * (size_t)C) &&
	          dl(out->support, s.d.con_support + lo, sizeof(int) * (size_t)C) &&
	          dl(out->uniq_support, s.d.con_uniq + lo, sizeof(int) * (size_t)C) &&
	          dl(out->max_read_len, s.d.con_maxlen + lo, sizeof(int) * (size_t)C) &&
	          dl(out->weight, s.d.con_weight + lo, sizeof(double) * (size_t)C) &&
	          dl(out->norm_abund, s.d.con_norm + lo, sizeof(double) * (size_t)C) &&
	          dl(out->pair_i, s.d.pair_i + lo, sizeof(int) * (size_t)M) && dl(out->pair_j, s.d.pair_j + lo, sizeof(int) * (size_t)M) &&
	          dl(out->pair_support, s.d.pair_support + lo, sizeof(int) * (size_t)M) &&
	          dl(out->pair_uniq, s.d.pair_uniq + lo, sizeof(int) * (size_t)M) &&
	          dl(out->pair_effective_count, s.d.pair_ec + lo, sizeof(int) * (size_t)M) &&
	          dl(out->pair_norm_abund, s.d.pair_norm + lo, sizeof(double) * (size_t)M);
	if (!ok) return ctx->fail("download failed");
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int psc_get_timing(psc_ctx *ctx, psc_timing *out)
{
	*out = ctx->timing;
	return 0;
}

} // extern "C"

#include "psc_api_solve.inl"
