// psc_api_solve.inl -- host orchestration of Solve: sizing, chunking (the GPU batch scheduler
// that replaces the worker pool of classes.cpp:318-477), kernel launches and result download.
// Included at the end of psc_api.cu.

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prologue(GraphDev g, SolveDev sv, int *fscratch, const int *fs_off, int *counter,
                                                  int width)
{
	Coop co = make_coop(width);
	for (;;) {
		int c = 0;
		if (co.lane == 0) c = atomic_add(counter, 1);
		c = co.bcast(c, 0);
		if (c >= g.n_comp) break;
		solve_prologue(co, g, sv, c, fscratch + fs_off[c]);
	}
}

// stack scratch of the enumeration: 3 ints per subexon of the component
__global__ void __launch_bounds__(128) k_enumerate(GraphDev g, SolveDev sv, const int *list, int nList, int *stk,
                                                   const int *fs_off, int *counter, int width)
{
	Coop co = make_coop(width);
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nList) break;
		int c = list[u];
		int n = g.comp_off[c + 1] - g.comp_off[c];
		int *b = stk + 3 * (long long)fs_off[c];
		enumerate_candidates(co, g, sv, c, b, b + n, b + 2 * n);
		co.sync();
	}
}

// Optional shared-memory slice per warp for the walk's frame memory.  Measured on B200 (full bench workload,
// total DP cycles / heaviest unit): 0 KB 1.78e12 / 2.73e9, 8 KB 1.83e12 / 3.27e9, 16 KB 1.81e12 / 3.08e9,
// 26 KB 2.07e12 / 3.31e9 -- since the walk keeps its hot state in registers, the L1 capacity the carve-out
// takes away costs more than the shorter latency gives.  So: off.
#ifndef PSC_DP_SMEM_PER_WARP
#define PSC_DP_SMEM_PER_WARP 0
#endif
#ifndef PSC_DP_MINBLOCKS
#define PSC_DP_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(128, PSC_DP_MINBLOCKS) k_dp_units(GraphDev g, SolveDev sv, const DpUnit *units, const int *order, int nUnits,
                                                  char *scratch, char *txArena, int *n_tx, int *counter, int width, int smemPerWarp, long long *dbg)
{
	Coop co = make_coop(width);
	char *sm = nullptr;
	size_t smBytes = 0;
#ifndef PSC_CPU_SHIM
	extern __shared__ __align__(16) unsigned char dp_smem[];
	if (width == 32 && smemPerWarp > 0) {
		sm = (char *)dp_smem + (size_t)(threadIdx.x >> 5) * smemPerWarp;
		smBytes = (size_t)smemPerWarp;
	}
#endif
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nUnits) break;
		int ui = order[u];
		dp_unit_pick(co, g, sv, units[ui], scratch, txArena, n_tx + ui, sm, smBytes, dbg ? dbg + 8 * (long long)ui : nullptr);
		co.sync();
	}
}

// The static-program DP engine (k_dp2.cuh): every cooperative group of the grid owns one global arena of arenaBytes
// and a slice of hotBytes of the block's dynamic shared memory, and pulls (component, sample) units, heaviest first.
// Units that outgrow their arena report n_tx = -1 and are counted in *overflow; the host redoes them with k_dp_units.
// The block size is 128 (four units side by side) for the bulk of the units and 32 for the heavy launch, whose single
// group gets most of an SM's shared memory.
#ifndef PSC_DP2_MINBLOCKS
#define PSC_DP2_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(128, PSC_DP2_MINBLOCKS) k_dp2_units(GraphDev g, SolveDev sv, const DpUnit *units, const int *order, int nUnits,
                                                                    char *arenas, size_t arenaBytes, size_t hotBytes, char *txArena, int *n_tx,
                                                                    int *counter, int *overflow, int width, long long *dbg, int roomy = 0)
{
	Coop co = make_coop(width);
	char *arena = arenas + (size_t)coop_id(width) * arenaBytes;
	char *hot = nullptr;
#ifndef PSC_CPU_SHIM
	extern __shared__ __align__(16) unsigned char dp2_smem[];
	if (width == 32 && hotBytes > 0) hot = (char *)dp2_smem + (size_t)(threadIdx.x >> 5) * hotBytes;
#endif
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nUnits) break;
		int ui = order[u];
		dp2_unit_pick(co, g, sv, units[ui], arena, arenaBytes, hot, hotBytes, txArena, n_tx + ui, dbg ? dbg + 8 * (long long)ui : nullptr, roomy);
		if (co.lane == 0 && n_tx[ui] < 0) atomic_add(overflow, 1);
		co.sync();
	}
}

#define PSC_MERGE_SORT_CAP 2048
// The same engine, launched over PACKETS: up to four units of similar cost whose shared-memory slices (each sized for
// its unit, dp2_smem_slice) fit one block's pool together.  Warp w of the block runs unit w of the packet; the block
// moves to the next packet when all of them are done.  A unit too big for the pool gets no slice (global arena).
struct DpPacket {
	int unit[4];     // index into units[], -1: none
	int hotOff[4];   // byte offset of the unit's slice in the block's pool
	int hotBytes[4]; // 0: no slice
};
#ifndef PSC_CPU_SHIM
__global__ void __launch_bounds__(128, PSC_DP2_MINBLOCKS) k_dp2_packets(GraphDev g, SolveDev sv, const DpUnit *units, const DpPacket *packets,
                                                                      int nPackets, char *arenas, size_t arenaBytes, char *txArena, int *n_tx,
                                                                      int *counter, int *overflow, long long *dbg)
{
	Coop co = make_coop(32);
	char *arena = arenas + (size_t)coop_id(32) * arenaBytes;
	extern __shared__ __align__(16) unsigned char dp2_pool[];
	__shared__ int s_packet;
	const int w = (int)(threadIdx.x >> 5);
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) s_packet = atomic_add(counter, 1);
		__syncthreads();
		const int pk = s_packet;
		if (pk >= nPackets) break;
		const DpPacket &P = packets[pk];
		const int ui = P.unit[w];
		if (ui >= 0) {
			char *hot = P.hotBytes[w] > 0 ? (char *)dp2_pool + P.hotOff[w] : nullptr;
			dp2_unit_pick(co, g, sv, units[ui], arena, arenaBytes, hot, (size_t)P.hotBytes[w], txArena, n_tx + ui, dbg ? dbg + 8 * (long long)ui : nullptr);
			if (co.lane == 0 && n_tx[ui] < 0) atomic_add(overflow, 1);
		}
	}
}
#endif

__global__ void __launch_bounds__(128) k_dp_merge(GraphDev g, SolveDev sv, const DpUnit *units, const int *compList, int nComp,
                                                  const int *n_tx, const char *txArena, char *scratch, const long long *scr_off,
                                                  int *counter, int width)
{
	Coop co = make_coop(width);
	int S = sv.prm.S;
	// sort buffer of CoalesceSameTranscripts: 2048 (prefix | index) words per warp in shared memory
	uint64_t *sort64 = nullptr;
	int sort64Cap = 0;
#ifndef PSC_CPU_SHIM
	// one warp per block and 16 KB of static shared memory: a kernel asking for 64 KB per block had to wait for SMs
	// to drain of the other chain's (shared-memory-free) kernels before it could start
	__shared__ __align__(16) uint64_t merge_smem[PSC_MERGE_SORT_CAP];
	if (width == 32) {
		sort64 = merge_smem;
		sort64Cap = PSC_MERGE_SORT_CAP;
	}
#else
	static thread_local uint64_t merge_host[PSC_MERGE_SORT_CAP];
	sort64 = merge_host;
	sort64Cap = PSC_MERGE_SORT_CAP;
#endif
	for (;;) {
		int k = 0;
		if (co.lane == 0) k = atomic_add(counter, 1);
		k = co.bcast(k, 0);
		if (k >= nComp) break;
		dp_merge(co, g, sv, compList[k], units, k * S, n_tx, txArena, scratch + scr_off[k], sort64, sort64Cap);
		co.sync();
	}
}

// phase: 1 pick round 1, 2 pick round 2 + filtration, 3 rescue.  list[0 .. nList): the chunk-local units of this pass,
// heaviest first
#ifndef PSC_UNITS_MINBLOCKS
#define PSC_UNITS_MINBLOCKS 4
#endif
// One kernel per phase (the phase is a template parameter): each phase's code is tens of thousands of instructions and
// the launches were instruction-fetch bound (ncu: `no_instruction` the top stall) while one kernel carried all three.
template <int PHASE>
__global__ void __launch_bounds__(128, PSC_UNITS_MINBLOCKS) k_units(GraphDev g, SolveDev sv, ChunkDev ch, const int *list, int nList,
                                                                   int *counter, int width, long long *dbg)
{
	Coop co = make_coop(width);
	int S = sv.prm.S;
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nList) break;
		int ul = list[u];
		int c = ch.c0 + ul / S, s = ul % S;
		long long t0 = dbg ? clock64() : 0;
		if (PHASE == 1) unit_pick_round1(co, g, sv, ch, c, s);
		else if (PHASE == 2) unit_pick_round2(co, g, sv, ch, c, s);
		else unit_rescue(co, g, sv, ch, c, s);
		co.sync();
		if (dbg && co.lane == 0) dbg[(long long)ul * 3 + (PHASE - 1)] = clock64() - t0; // PSC_DEBUG_UNITS
	}
}

__global__ void __launch_bounds__(128) k_emit(GraphDev g, SolveDev sv, ChunkDev ch, OutDev o, const long long *tx_off,
                                              const long long *ex_off, const int *list, int nList, int *counter, int width)
{
	Coop co = make_coop(width);
	int S = sv.prm.S;
	for (;;) {
		int u = 0;
		if (co.lane == 0) u = atomic_add(counter, 1);
		u = co.bcast(u, 0);
		if (u >= nList) break;
		int ul = list[u];
		int c = ch.c0 + ul / S, s = ul % S;
		unit_emit(co, g, sv, ch, c, s, o, tx_off[ul], ex_off[ul]);
		co.sync();
	}
}

// psc_solve runs the per-unit kernels as two passes over disjoint sets of components (the bulk, the heavy DP
// components); each pass delivers its transcripts in component order.  Merge the two lists into the single
// component-major order the C ABI promises: one pass over the data, split over a few host threads.
static void merge_passes(ResultSet &out, const ResultSet &A, const ResultSet &B)
{
	const size_t nA = A.r_comp.size(), nB = B.r_comp.size(), nt = nA + nB;
	// source of output position k: index into A, or nA + index into B
	std::vector<uint32_t> perm(nt);
	{
		size_t a = 0, b = 0, k = 0;
		while (a < nA && b < nB) perm[k++] = A.r_comp[a] < B.r_comp[b] ? (uint32_t)a++ : (uint32_t)(nA + b++);
		while (a < nA) perm[k++] = (uint32_t)a++;
		while (b < nB) perm[k++] = (uint32_t)(nA + b++);
	}
	out.r_exon_off.resize(nt + 1);
	{
		long long e = 0;
		for (size_t k = 0; k < nt; ++k) {
			out.r_exon_off[k] = e;
			size_t p = perm[k];
			e += p < nA ? A.r_exon_off[p + 1] - A.r_exon_off[p] : B.r_exon_off[p - nA + 1] - B.r_exon_off[p - nA];
		}
		out.r_exon_off[nt] = e;
	}
	const size_t ne = (size_t)out.r_exon_off[nt];
	out.r_comp.resize(nt); out.r_sample.resize(nt); out.r_chr.resize(nt); out.r_gene.resize(nt); out.r_strand.resize(nt);
	out.r_fpkm.resize(nt); out.r_cov.resize(nt); out.r_exon_start.resize(ne); out.r_exon_end.resize(ne);
	auto work = [&](size_t k0, size_t k1) {
		for (size_t k = k0; k < k1; ++k) {
			size_t p = perm[k];
			const ResultSet &R = p < nA ? A : B;
			if (p >= nA) p -= nA;
			out.r_comp[k] = R.r_comp[p];
			out.r_sample[k] = R.r_sample[p];
			out.r_chr[k] = R.r_chr[p];
			out.r_gene[k] = R.r_gene[p];
			out.r_strand[k] = R.r_strand[p];
			out.r_fpkm[k] = R.r_fpkm[p];
			out.r_cov[k] = R.r_cov[p];
			const long long s0 = R.r_exon_off[p], n = R.r_exon_off[p + 1] - s0, e = out.r_exon_off[k];
			if (n > 0) {
				memcpy(&out.r_exon_start[(size_t)e], &R.r_exon_start[(size_t)s0], sizeof(int) * (size_t)n);
				memcpy(&out.r_exon_end[(size_t)e], &R.r_exon_end[(size_t)s0], sizeof(int) * (size_t)n);
			}
		}
	};
	size_t nThreads = nt >= 65536 ? 4 : 1;
	if (const char *e = getenv("PSC_MERGE_THREADS")) nThreads = (size_t)std::max(1, atoi(e)); // tests
	if (nThreads > nt) nThreads = 1;
	std::vector<std::thread> pool;
	for (size_t t = 1; t < nThreads; ++t) pool.emplace_back(work, nt * t / nThreads, nt * (t + 1) / nThreads);
	work(0, nt / nThreads);
	for (auto &th : pool) th.join();
}

static int launch_grid(psc_ctx *ctx, long long units)
{
	long long groupsPerBlock = 128 / ctx->coop;
	long long blocks = (units + groupsPerBlock - 1) / groupsPerBlock;
	long long maxBlocks = (long long)ctx->n_sm * 16;
	if (blocks > maxBlocks) blocks = maxBlocks;
	if (blocks < 1) blocks = 1;
	return (int)blocks;
}

extern "C" {

int psc_solve(psc_ctx *ctx)
{
	cudaSetDevice(ctx->device);
	if (!ctx->constraints_built) return ctx->fail("psc_build_constraints must run before psc_solve");
	SolveState &st = ctx->solve;
	const int S = ctx->prm.n_samples, nc = ctx->gd.n_comp, ns = ctx->gd.n_se;
	std::atomic<long long> launches{0};
	// PSC_DEBUG_HOST=1: wall-clock of the host-side sections (stderr)
	const bool hostDbg = getenv("PSC_DEBUG_HOST") != nullptr;
	const auto solveT0 = std::chrono::steady_clock::now();
	const std::thread::id mainThread = std::this_thread::get_id();
	std::mutex lapMu;
	std::vector<std::pair<std::thread::id, std::chrono::steady_clock::time_point>> lapT0; // per host thread
	auto lap = [&](const char *what) {
		if (!hostDbg) return;
		std::lock_guard<std::mutex> lk(lapMu);
		const std::thread::id me = std::this_thread::get_id();
		size_t who = 0;
		while (who < lapT0.size() && lapT0[who].first != me) ++who;
		if (who == lapT0.size()) lapT0.push_back({me, solveT0});
		auto t = std::chrono::steady_clock::now();
		fprintf(stderr, "[host%s] %-28s %8.2f ms  (at %8.2f)\n", me == mainThread ? "" : (who == 1 ? " 2" : " 3"), what,
		        std::chrono::duration<double, std::milli>(t - lapT0[who].second).count(), std::chrono::duration<double, std::milli>(t - solveT0).count());
		lapT0[who].second = t;
	};
	CK(cudaStreamSynchronize(ctx->stream)); // uploads and the constraint build ran on the context's primary stream
	if (!getenv("PSC_MEM_BUDGET_MB")) {
		// scratch budget per buffer class, from what is free NOW (reads and tables are resident): up to five such
		// buffers are alive at a time (unit scratch of the two chains, DP arenas of the two chains, transcript lists)
		size_t fr = 0, tot = 0;
		if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
			size_t held = st.ub[0].scratch.bytes + st.ub[1].scratch.bytes + st.db[0].dp_arena.bytes + st.db[1].dp_arena.bytes; // reusable
			ctx->mem_budget = std::max<size_t>((size_t)256 << 20, (fr + held) / 5);
		}
	}
	CK(cudaEventRecord(ctx->ev0, ctx->sMain));
	st.reset();
	st.h_at_cnt.assign((size_t)nc, 0); st.h_use_dp.assign((size_t)nc, 0); st.h_n_cand.assign((size_t)nc, 0);
	if (nc == 0) {
		ctx->solved = true;
		return 0;
	}
	int nEdges = 0;
	for (int c = 0; c < nc; ++c) nEdges += ctx->h_next_cnt[(size_t)c];

	SolveDev sv;
	memset(&sv, 0, sizeof(sv));
	sv.prm.fpkmFraction = ctx->prm.fpkm_fraction;
	sv.prm.classifierThreshold = ctx->prm.classifier_threshold;
	sv.prm.minDepth = ctx->prm.min_read_depth;
	sv.prm.maxDpConstraintSize = ctx->prm.max_dp_constraint_size;
	sv.prm.S = S;
	sv.prm.readLen = ctx->prm.read_len;
	sv.prm.fragLen = ctx->prm.frag_len;
	sv.prm.fragStdev = ctx->prm.frag_stdev;
	sv.prm.matePaired = ctx->prm.mate_paired;
	sv.samples = ctx->d_samples.as<SampleDev>();

	std::vector<int> fsOff((size_t)nc + 1);
	for (int c = 0; c <= nc; ++c) fsOff[(size_t)c] = ctx->h_comp_off[(size_t)c];
	bool ok = st.can_start.ensure((size_t)ns + 1) && st.can_end.ensure((size_t)ns + 1) &&
	          st.edge_support.ensure(sizeof(int) * ((size_t)nEdges + 1)) && st.cs00.ensure(sizeof(int) * (size_t)nc) &&
	          st.at_cnt.ensure(sizeof(int) * (size_t)nc) && st.use_dp.ensure(sizeof(int) * (size_t)nc) &&
	          st.def_gene.ensure(sizeof(int) * 2 * (size_t)nc) && st.n_cand.ensure(sizeof(int) * (size_t)nc) &&
	          st.error_flag.ensure(64) && st.fscratch.ensure(sizeof(int) * ((size_t)ns + 1)) &&
	          upload(ctx, st.fs_off, fsOff.data(), (size_t)nc + 1);
	if (!ok) return ctx->fail("out of device memory (solve state)");
	CK(cudaMemsetAsync(st.error_flag.p, 0, 64, ctx->sMain));
	sv.can_start = st.can_start.as<uint8_t>();
	sv.can_end = st.can_end.as<uint8_t>();
	sv.edge_support = st.edge_support.as<int>();
	sv.cs00 = st.cs00.as<int>();
	sv.at_cnt = st.at_cnt.as<int>();
	sv.use_dp = st.use_dp.as<int>();
	sv.def_gene = st.def_gene.as<int>();
	sv.n_cand = st.n_cand.as<int>();
	sv.error_flag = st.error_flag.as<int>();
	int *counter = ctx->d_counter.as<int>();

	// the edge_support array is indexed with the global next-edge index: next_off is global already
	// ---- P1: prologue ----
	CK(cudaMemsetAsync(counter, 0, 64, ctx->sMain));
	KLAUNCH_ON(ctx->sMain, 2, k_prologue, launch_grid(ctx, nc), 128, ctx->gd, sv, st.fscratch.as<int>(), st.fs_off.as<int>(), counter, ctx->coop);
	++launches;
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(st.h_at_cnt.data(), sv.at_cnt, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->sMain));
	CK(cudaMemcpyAsync(st.h_use_dp.data(), sv.use_dp, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->sMain));
	CK(cudaStreamSynchronize(ctx->sMain));

	lap("prologue");
	// ---- candidate capacities ----
	std::vector<long long> rowOff((size_t)nc + 1), bitsOff((size_t)nc + 1);
	std::vector<int> enumList, dpList;
	{
		long long r = 0, b = 0;
		for (int c = 0; c < nc; ++c) {
			int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
			int W = (n + 63) / 64;
			long long cap;
			if (st.h_use_dp[(size_t)c]) {
				int Cmax = 0;
				for (int s = 0; s < S; ++s) Cmax = std::max(Cmax, ctx->samples[(size_t)s].h_ncon[(size_t)c]);
				cap = 1000 + (long long)Cmax + 100LL * (S - 1) + 8;
				dpList.push_back(c);
			} else {
				cap = std::max(0, st.h_at_cnt[(size_t)c]);
				enumList.push_back(c);
			}
			rowOff[(size_t)c] = r;
			bitsOff[(size_t)c] = b;
			r += cap;
			b += cap * W;
		}
		rowOff[(size_t)nc] = r;
		bitsOff[(size_t)nc] = b;
		ok = upload(ctx, st.cand_row_off, rowOff.data(), (size_t)nc + 1) && upload(ctx, st.cand_bits_off, bitsOff.data(), (size_t)nc + 1) &&
		     st.cand_bits.ensure(sizeof(uint64_t) * (size_t)(b + 1)) && st.cand_first.ensure(sizeof(int) * (size_t)(r + 1)) &&
		     st.cand_last.ensure(sizeof(int) * (size_t)(r + 1)) && st.cand_corr.ensure((size_t)(r + 1)) &&
		     st.cand_sup1.ensure(sizeof(int) * (size_t)(r + 1)) && st.cand_sup2.ensure(sizeof(int) * (size_t)(r + 1));
		if (!ok) return ctx->fail("out of device memory (candidate transcripts)");
	}
	sv.cand_row_off = st.cand_row_off.as<long long>();
	sv.cand_bits_off = st.cand_bits_off.as<long long>();
	sv.cand_bits = st.cand_bits.as<uint64_t>();
	sv.cand_first = st.cand_first.as<int>();
	sv.cand_last = st.cand_last.as<int>();
	sv.cand_corr = st.cand_corr.as<int8_t>();
	sv.cand_sup1 = st.cand_sup1.as<int>();
	sv.cand_sup2 = st.cand_sup2.as<int>();

	lap("candidate capacities");
	// ---- P2a: enumeration ----
	if (!enumList.empty()) {
		if (!upload(ctx, st.list, enumList.data(), enumList.size()) || !st.stk.ensure(sizeof(int) * 3 * ((size_t)ns + 1)))
			return ctx->fail("out of device memory");
		CK(cudaMemsetAsync(counter, 0, 64, ctx->sMain));
		// the enumeration is sequential per component: always one thread per unit
		int saved = ctx->coop;
		ctx->coop = 1;
		int grid = launch_grid(ctx, (long long)enumList.size());
		ctx->coop = saved;
		KLAUNCH_ON(ctx->sMain, 3, k_enumerate, grid, 128, ctx->gd, sv, st.list.as<int>(), (int)enumList.size(), st.stk.as<int>(), st.fs_off.as<int>(),
		        counter, 1);
		++launches;
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(ctx->sMain)); // enumList staging buffer is reused below
	}
	CK(cudaMemcpyAsync(st.h_n_cand.data(), sv.n_cand, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->sMain));
	CK(cudaStreamSynchronize(ctx->sMain));
	// ---- P3..P6 per chunk of components, as a pass over all components (pass 0), over all but the heavy DP
	// components (1) or over the heavy DP components (2).  Pass 1 runs on the second stream with its own buffers
	// while the heavy DP units are still busy on the main one. ----
	std::vector<char> group((size_t)nc, 0); // 0: enumerated component, 1: DP component of the bulk, 2: heavy DP component
	// pass 0: every component; 1: everything but the heavy ones; 2: the heavy ones; 3: the enumerated ones; 4: the bulk DP ones
	auto in_pass = [&](int pass, int c) {
		const int gr = group[(size_t)c];
		return pass == 0 || (pass == 1 && gr != 2) || (pass == 2 && gr == 2) || (pass == 3 && gr == 0) || (pass == 4 && gr == 1);
	};
	auto run_units = [&](int pass, cudaStream_t strm, UnitBufs &ub, int *cnt, ResultSet &R) -> int {
		// per (component, sample) scratch needs of this pass, computed on a few host threads
		std::vector<long long> soAll((size_t)nc * S, 0), poAll((size_t)nc * S, 0), sNeedAll((size_t)nc, 0), pNeedAll((size_t)nc, 0);
		{
			auto work = [&](int ca, int cb) {
				for (int c = ca; c < cb; ++c) {
					const bool mine = in_pass(pass, c);
					if (!mine) continue;
					int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
					int A = st.h_n_cand[(size_t)c];
					long long sNeed = 0, pNeed = 0;
					for (int s = 0; s < S; ++s) {
						int M = ctx->samples[(size_t)s].h_npair[(size_t)c], C = ctx->samples[(size_t)s].h_ncon[(size_t)c];
						int cap = std::max(1, std::min(A, M + 1));
						soAll[(size_t)c * S + s] = sNeed;
						poAll[(size_t)c * S + s] = pNeed;
						sNeed += (long long)unit_scratch_layout(nullptr, nullptr, A, M, cap, n, C, ctx->h_comp_ng[(size_t)c], ctx->h_next_cnt[(size_t)c]);
						pNeed += (long long)unit_pred_layout(nullptr, nullptr, cap);
					}
					sNeedAll[(size_t)c] = sNeed;
					pNeedAll[(size_t)c] = pNeed;
				}
			};
			const int nThreads = nc >= 8192 ? 4 : 1;
			std::vector<std::thread> pool;
			for (int t = 1; t < nThreads; ++t) pool.emplace_back(work, (int)((long long)nc * t / nThreads), (int)((long long)nc * (t + 1) / nThreads));
			work(0, nc / nThreads);
			for (auto &th : pool) th.join();
		}
		int c0 = 0;
		while (c0 < nc) {
			std::vector<long long> scrOff, predOff;
			std::vector<long long> cost;
			std::vector<int> order; // chunk-local ids of the units of this pass
			long long sAcc = 0, pAcc = 0;
			int c1 = c0;
			while (c1 < nc) {
				int A = st.h_n_cand[(size_t)c1];
				const long long sNeed = sNeedAll[(size_t)c1], pNeed = pNeedAll[(size_t)c1];
				const long long *so = &soAll[(size_t)c1 * S], *po = &poAll[(size_t)c1 * S];
				const bool mine = in_pass(pass, c1);
				if (c1 > c0 && (size_t)(sAcc + sNeed) > ctx->mem_budget) break;
				for (int s = 0; s < S; ++s) {
					scrOff.push_back(sAcc + so[(size_t)s]);
					predOff.push_back(pAcc + po[(size_t)s]);
					int M = ctx->samples[(size_t)s].h_npair[(size_t)c1];
					cost.push_back((long long)A * (M + 1));
					if (mine) order.push_back((c1 - c0) * S + s);
				}
				sAcc += sNeed;
				pAcc += pNeed;
				++c1;
			}
			scrOff.push_back(sAcc);
			predOff.push_back(pAcc);
			int nUnits = (c1 - c0) * S;
			int nList = (int)order.size();
			if (nList == 0) {
				c0 = c1;
				continue;
			}
			if (order.size() > 65536) {
				// heaviest first is only a scheduling hint: bucket by the bit length of the cost instead of a full sort
				std::vector<int> bucketOf(order.size());
				size_t cnt64[66] = {0};
				for (size_t q = 0; q < order.size(); ++q) {
					long long cq = cost[(size_t)order[q]];
					int b = 0;
					while (cq > 0) { cq >>= 1; ++b; }
					bucketOf[q] = 64 - b; // 0: heaviest
					++cnt64[bucketOf[q] + 1];
				}
				for (int b = 0; b < 65; ++b) cnt64[b + 1] += cnt64[b];
				std::vector<int> sorted(order.size());
				for (size_t q = 0; q < order.size(); ++q) sorted[cnt64[bucketOf[q]]++] = order[q];
				order.swap(sorted);
			} else
				std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
			ok = ub.scratch.ensure((size_t)sAcc + 16) && ub.pred.ensure((size_t)pAcc + 16) &&
			     upload(ctx, ub.scr_off, scrOff.data(), scrOff.size(), strm) && upload(ctx, ub.pred_off, predOff.data(), predOff.size(), strm) &&
			     upload(ctx, ub.order, order.data(), order.size(), strm) && ub.n_pred.ensure(sizeof(int) * (size_t)nUnits) &&
			     ub.n_raw.ensure(sizeof(int) * (size_t)nUnits) && ub.n_exon.ensure(sizeof(int) * (size_t)nUnits);
			if (!ok) return ctx->fail("out of device memory (unit scratch)");
			lap("units host prep");
			ChunkDev ch;
			ch.c0 = c0;
			ch.c1 = c1;
			ch.scratch_off = ub.scr_off.as<long long>();
			ch.scratch = (char *)ub.scratch.p;
			ch.pred_off = ub.pred_off.as<long long>();
			ch.pred = (char *)ub.pred.p;
			ch.n_pred = ub.n_pred.as<int>();
			ch.n_raw = ub.n_raw.as<int>();
			ch.n_exon = ub.n_exon.as<int>();
			CK(cudaMemsetAsync(ch.n_pred, 0, sizeof(int) * (size_t)nUnits, strm));
			CK(cudaMemsetAsync(ch.n_raw, 0, sizeof(int) * (size_t)nUnits, strm));
			CK(cudaMemsetAsync(ch.n_exon, 0, sizeof(int) * (size_t)nUnits, strm));
			long long *unitDbg = nullptr;
			if (getenv("PSC_DEBUG_UNITS")) {
				if (!ub.dbg.ensure(sizeof(long long) * 3 * (size_t)(nUnits + 1))) return ctx->fail("out of device memory");
				CK(cudaMemsetAsync(ub.dbg.p, 0, sizeof(long long) * 3 * (size_t)(nUnits + 1), strm));
				unitDbg = ub.dbg.as<long long>();
			}
			for (int phase = 1; phase <= 3; ++phase) {
				CK(cudaMemsetAsync(cnt, 0, 32, strm));
				if (phase == 1) KLAUNCH_ON(strm, 5, k_units<1>, launch_grid(ctx, nList), 128, ctx->gd, sv, ch, ub.order.as<int>(), nList, cnt, ctx->coop, unitDbg);
				else if (phase == 2) KLAUNCH_ON(strm, 8, k_units<2>, launch_grid(ctx, nList), 128, ctx->gd, sv, ch, ub.order.as<int>(), nList, cnt, ctx->coop, unitDbg);
				else KLAUNCH_ON(strm, 9, k_units<3>, launch_grid(ctx, nList), 128, ctx->gd, sv, ch, ub.order.as<int>(), nList, cnt, ctx->coop, unitDbg);
				++launches;
				CK(cudaGetLastError());
				// One launch in flight per chain: with the three phases queued back to back the two chains' kernels
				// serialised behind each other in the hardware queues (measured: 800 ms per step instead of 640).
				CK(cudaStreamSynchronize(strm));
				lap(phase == 1 ? "  units phase 1 (pick)" : (phase == 2 ? "  units phase 2 (pick, refine, EM)" : "  units phase 3 (rescue)"));
			}
			if (unitDbg) { // PSC_DEBUG_UNITS=1: the slowest (component, sample) units of this pass
				CK(cudaStreamSynchronize(strm));
				std::vector<long long> h((size_t)nUnits * 3);
				cudaMemcpy(h.data(), unitDbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
				for (int ph = 0; ph < 2; ++ph) {
					std::vector<int> ix(order);
					std::sort(ix.begin(), ix.end(), [&](int a, int b) { return h[(size_t)a * 3 + ph] > h[(size_t)b * 3 + ph]; });
					long long tot = 0;
					for (int u : order) tot += h[(size_t)u * 3 + ph];
					fprintf(stderr, "[units] pass %d phase %d: %d units, total %.1f Mcycles\n", pass, ph + 1, nList, tot / 1e6);
					for (int k = 0; k < 5 && k < nList; ++k) {
						int u = ix[(size_t)k], c = c0 + u / S, sm = u % S;
						fprintf(stderr, "[units]   comp %d sample %d n=%d A=%d C=%d M=%d: %.1f Mcycles\n", c, sm,
						        ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c], st.h_n_cand[(size_t)c],
						        ctx->samples[(size_t)sm].h_ncon[(size_t)c], ctx->samples[(size_t)sm].h_npair[(size_t)c], h[(size_t)u * 3 + ph] / 1e6);
					}
				}
			}
			std::vector<int> hPred((size_t)nUnits), hExon((size_t)nUnits);
			CK(cudaMemcpyAsync(hPred.data(), ch.n_pred, sizeof(int) * (size_t)nUnits, cudaMemcpyDeviceToHost, strm));
			CK(cudaMemcpyAsync(hExon.data(), ch.n_exon, sizeof(int) * (size_t)nUnits, cudaMemcpyDeviceToHost, strm));
			CK(cudaStreamSynchronize(strm));
			lap("units kernels");
			std::vector<long long> txOff((size_t)nUnits + 1), exOff((size_t)nUnits + 1);
			long long nt = 0, ne = 0;
			for (int u = 0; u < nUnits; ++u) {
				txOff[(size_t)u] = nt;
				exOff[(size_t)u] = ne;
				nt += hPred[(size_t)u];
				ne += hExon[(size_t)u];
			}
			txOff[(size_t)nUnits] = nt;
			exOff[(size_t)nUnits] = ne;
			if (nt > 0) {
				ok = upload(ctx, ub.tx_off, txOff.data(), txOff.size(), strm) && upload(ctx, ub.ex_off, exOff.data(), exOff.size(), strm) &&
				     ub.o_comp.ensure(sizeof(int) * (size_t)nt) && ub.o_sample.ensure(sizeof(int) * (size_t)nt) &&
				     ub.o_chr.ensure(sizeof(int) * (size_t)nt) && ub.o_gene.ensure(sizeof(int) * (size_t)nt) &&
				     ub.o_strand.ensure((size_t)nt) && ub.o_exon_off.ensure(sizeof(long long) * (size_t)nt) &&
				     ub.o_exon_start.ensure(sizeof(int) * (size_t)(ne + 1)) && ub.o_exon_end.ensure(sizeof(int) * (size_t)(ne + 1)) &&
				     ub.o_fpkm.ensure(sizeof(double) * (size_t)nt) && ub.o_cov.ensure(sizeof(double) * (size_t)nt);
				if (!ok) return ctx->fail("out of device memory (output)");
				OutDev o;
				o.comp = ub.o_comp.as<int>();
				o.sample = ub.o_sample.as<int>();
				o.chr = ub.o_chr.as<int>();
				o.gene = ub.o_gene.as<int>();
				o.strand = ub.o_strand.as<int8_t>();
				o.exon_off = ub.o_exon_off.as<long long>();
				o.exon_start = ub.o_exon_start.as<int>();
				o.exon_end = ub.o_exon_end.as<int>();
				o.fpkm = ub.o_fpkm.as<double>();
				o.cov = ub.o_cov.as<double>();
				lap("emit host prep");
				CK(cudaMemsetAsync(cnt, 0, 32, strm));
				KLAUNCH_ON(strm, 6, k_emit, launch_grid(ctx, nList), 128, ctx->gd, sv, ch, o, ub.tx_off.as<long long>(), ub.ex_off.as<long long>(),
				           ub.order.as<int>(), nList, cnt, ctx->coop);
				++launches;
				CK(cudaGetLastError());
				size_t t0 = R.r_comp.size(), e0 = R.r_exon_start.size();
				R.r_comp.resize(t0 + (size_t)nt); R.r_sample.resize(t0 + (size_t)nt); R.r_chr.resize(t0 + (size_t)nt);
				R.r_gene.resize(t0 + (size_t)nt); R.r_strand.resize(t0 + (size_t)nt); R.r_fpkm.resize(t0 + (size_t)nt);
				R.r_cov.resize(t0 + (size_t)nt); R.r_exon_start.resize(e0 + (size_t)ne); R.r_exon_end.resize(e0 + (size_t)ne);
				// device -> pinned staging -> result vectors
				const size_t NT = (size_t)nt, NE = (size_t)ne;
				const size_t oComp = 0, oSample = oComp + 4 * NT, oChr = oSample + 4 * NT, oGene = oChr + 4 * NT, oFpkm = (oGene + 4 * NT + 7) & ~(size_t)7,
				             oCov = oFpkm + 8 * NT, oEo = oCov + 8 * NT, oEs = oEo + 8 * NT, oEe = oEs + 4 * NE, oStrand = oEe + 4 * NE, total = oStrand + NT;
				if (!ub.stage.ensure(total + 64)) return ctx->fail("out of pinned host memory (result staging)");
				char *hs = (char *)ub.stage.p;
				CK(cudaMemcpyAsync(hs + oComp, o.comp, 4 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oSample, o.sample, 4 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oChr, o.chr, 4 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oGene, o.gene, 4 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oStrand, o.strand, NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oFpkm, o.fpkm, 8 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oCov, o.cov, 8 * NT, cudaMemcpyDeviceToHost, strm));
				CK(cudaMemcpyAsync(hs + oEo, o.exon_off, 8 * NT, cudaMemcpyDeviceToHost, strm));
				if (NE) {
					CK(cudaMemcpyAsync(hs + oEs, o.exon_start, 4 * NE, cudaMemcpyDeviceToHost, strm));
					CK(cudaMemcpyAsync(hs + oEe, o.exon_end, 4 * NE, cudaMemcpyDeviceToHost, strm));
				}
				CK(cudaStreamSynchronize(strm));
				memcpy(R.r_comp.data() + t0, hs + oComp, 4 * NT);
				memcpy(R.r_sample.data() + t0, hs + oSample, 4 * NT);
				memcpy(R.r_chr.data() + t0, hs + oChr, 4 * NT);
				memcpy(R.r_gene.data() + t0, hs + oGene, 4 * NT);
				memcpy(R.r_strand.data() + t0, hs + oStrand, NT);
				memcpy(R.r_fpkm.data() + t0, hs + oFpkm, 8 * NT);
				memcpy(R.r_cov.data() + t0, hs + oCov, 8 * NT);
				if (NE) {
					memcpy(R.r_exon_start.data() + e0, hs + oEs, 4 * NE);
					memcpy(R.r_exon_end.data() + e0, hs + oEe, 4 * NE);
				}
				const long long *eo = (const long long *)(hs + oEo);
				lap("emit + copy back");
				R.r_exon_off.pop_back();
				R.r_exon_off.reserve(R.r_exon_off.size() + NT + 1);
				for (long long k = 0; k < nt; ++k) R.r_exon_off.push_back((long long)e0 + eo[(size_t)k]);
				R.r_exon_off.push_back((long long)e0 + ne);
			}
			c0 = c1;
		}
		return 0;
	};
	lap("enumerate");
	// ---- P2b: DP.  Units are (component, sample); chunks of components that fit the memory budget ----
	long long dpBytes = 0;
	// One chunk of DP components in flight: dp_launch issues the static-program kernel asynchronously, dp_finish waits
	// for it, redoes the units that overflowed their arena with the walk engine and launches the merge.
	struct DpPending {
		bool active = false;
		std::vector<DpUnit> units;
		std::vector<int> comps;
		std::vector<long long> mergeOff;
		long long mAcc = 0;
		int nComp = 0;
	};
	const bool dpDbg = getenv("PSC_DEBUG_DP") != nullptr;
	const bool useV2 = !(getenv("PSC_DP_V2") && atoi(getenv("PSC_DP_V2")) == 0);
	auto dp_print = [&](const std::vector<DpUnit> &units, cudaStream_t strm, const char *what) {
		int nUnits = (int)units.size();
		CK(cudaStreamSynchronize(strm));
		std::vector<long long> h((size_t)nUnits * 8);
		cudaMemcpy(h.data(), st.dp_dbg.p, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
		std::vector<int> ix((size_t)nUnits);
		for (int u = 0; u < nUnits; ++u) ix[(size_t)u] = u;
		std::sort(ix.begin(), ix.end(), [&](int a, int b) { return h[(size_t)a * 8] > h[(size_t)b * 8]; });
		long long tot = 0, totStates = 0, totFrames = 0, maxKeys = 0, maxFrames = 0, maxList = 0, nOvf = 0;
		for (int u = 0; u < nUnits; ++u) {
			tot += h[(size_t)u * 8];
			totStates += h[(size_t)u * 8 + 1] & 0xffffffff;
			totFrames += (h[(size_t)u * 8 + 3] >> 12) & 0xfffffff;
			maxKeys = std::max(maxKeys, h[(size_t)u * 8 + 4]);
			maxFrames = std::max(maxFrames, h[(size_t)u * 8 + 5]);
			maxList = std::max(maxList, h[(size_t)u * 8 + 6]);
			if (h[(size_t)u * 8 + 7] < 0) ++nOvf;
		}
		fprintf(stderr, "[dp %s] %d units, total %.1f Mcycles, %lld states, %lld frames; program tables max keys %lld frames %lld list %lld, %lld overflowed\n",
		        what, nUnits, tot / 1e6, totStates, totFrames, maxKeys, maxFrames, maxList, nOvf);
		for (int k = 0; k < 12 && k < nUnits; ++k) {
			int u = ix[(size_t)k];
			int c = units[(size_t)u].comp, s = units[(size_t)u].sample;
			fprintf(stderr, "[dp %s] unit %d comp %d sample %d n=%d C=%d iterBound=%d: %.1f Mcycles, states %lld levels %lld iters %lld frames %lld roots %lld | Mcyc in roots %lld, in assign %lld | keys %lld frames %lld list %lld cap %lld\n", what, u, c, s,
			        ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c], ctx->samples[(size_t)s].h_ncon[(size_t)c], units[(size_t)u].iterBound,
			        h[(size_t)u * 8] / 1e6, h[(size_t)u * 8 + 1] & 0xffffffff, h[(size_t)u * 8 + 2] & 0xffffffff, h[(size_t)u * 8 + 3] & 0xfff, (h[(size_t)u * 8 + 3] >> 12) & 0xfffffff, h[(size_t)u * 8 + 3] >> 40,
			        h[(size_t)u * 8 + 1] >> 32, h[(size_t)u * 8 + 2] >> 32, h[(size_t)u * 8 + 4], h[(size_t)u * 8 + 5], h[(size_t)u * 8 + 6], h[(size_t)u * 8 + 7]);
		}
		return 0;
	};
	// redo the listed units with the walk engine, in sub-chunks whose scratch fits the budget
	auto dp_walk_units = [&](DpPending &pd, const std::vector<int> &redo, cudaStream_t strm, DpBufs &db, int *cnt) -> int {
		size_t k0 = 0;
		while (k0 < redo.size()) {
			long long sAcc = 0;
			size_t k1 = k0;
			std::vector<int> order;
			std::vector<long long> cost;
			while (k1 < redo.size()) {
				DpUnit &u = pd.units[(size_t)redo[k1]];
				int c = u.comp;
				int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
				if (n > 10000) return ctx->fail("component with more than 10000 subexons overflowed the DP program arena (raise PSC_MEM_BUDGET_MB)");
				int C = ctx->samples[(size_t)u.sample].h_ncon[(size_t)c];
				long long need = (long long)dp_scratch_layout(nullptr, nullptr, n, ctx->h_next_cnt[(size_t)c], std::max(C, 100), 1);
				if (k1 > k0 && (size_t)(sAcc + need) > ctx->mem_budget) break;
				u.scratch_off = sAcc;
				sAcc += need;
				order.push_back(redo[k1]);
				cost.push_back((long long)u.iterBound * (C + 1));
				++k1;
			}
			std::vector<int> ix(order.size());
			for (size_t q = 0; q < ix.size(); ++q) ix[q] = (int)q;
			std::stable_sort(ix.begin(), ix.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
			std::vector<int> sorted(order.size());
			for (size_t q = 0; q < ix.size(); ++q) sorted[q] = order[(size_t)ix[q]];
			if (!(*db.scratch).ensure((size_t)sAcc + 16) || !upload(ctx, db.dp_units, pd.units.data(), pd.units.size(), strm) ||
			    !upload(ctx, db.dp_order, sorted.data(), sorted.size(), strm))
				return ctx->fail("out of device memory (DP scratch)");
			CK(cudaMemsetAsync(cnt, 0, 32, strm));
			const int nRedo = (int)sorted.size();
			const int smemPerWarp = ctx->coop == 32 ? PSC_DP_SMEM_PER_WARP : 0;
			long long *dbgPtr = dpDbg ? st.dp_dbg.as<long long>() : nullptr;
#ifndef PSC_CPU_SHIM
			CK(cudaFuncSetAttribute(k_dp_units, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * PSC_DP_SMEM_PER_WARP));
			{
			std::lock_guard<std::mutex> lk_(ctx->launchMu);
			ctx->ktime_begin(4, strm);
			k_dp_units<<<launch_grid(ctx, nRedo), 128, 4 * smemPerWarp, strm>>>(
			    ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>(), nRedo, (char *)(*db.scratch).p, (char *)db.dp_tx.p,
			    db.dp_ntx.as<int>(), cnt, ctx->coop, smemPerWarp, dbgPtr);
			ctx->ktime_end();
			}
#else
			KLAUNCH_ON(strm, 4, k_dp_units, launch_grid(ctx, nRedo), 128, ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>(), nRedo,
			           (char *)(*db.scratch).p, (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt, ctx->coop, smemPerWarp, dbgPtr);
#endif
			++launches;
			CK(cudaGetLastError());
			CK(cudaStreamSynchronize(strm)); // the scratch and the order list are reused by the next sub-chunk
			k0 = k1;
		}
		return 0;
	};
	// second attempt of the units that outgrew their arena: the program engine again, one roomy arena per group
	// (four times the keys, 24 frames per key), in sub-chunks that fit the budget.  Returns the units still overflowing.
	auto dp_program_retry = [&](DpPending &pd, std::vector<int> &redo, cudaStream_t strm, DpBufs &db, int *cnt) -> int {
		std::vector<int> still;
		size_t k0 = 0;
		const int groupsPerBlock = 128 / ctx->coop;
		while (k0 < redo.size()) {
			size_t arenaR = 0, k1 = k0;
			while (k1 < redo.size()) {
				const DpUnit &u = pd.units[(size_t)redo[k1]];
				int n = ctx->h_comp_off[(size_t)u.comp + 1] - ctx->h_comp_off[(size_t)u.comp];
				int C = ctx->samples[(size_t)u.sample].h_ncon[(size_t)u.comp];
				long long keys = std::min<long long>(4LL * dp2_wanted_keys(n), 262144);
				size_t need = dp2_layout(nullptr, nullptr, nullptr, 0, nullptr, 0, n, ctx->h_next_cnt[(size_t)u.comp], std::max(C, 100), (int)keys, DP2_FR_PER_KEY_ROOMY);
				need = (need + 255) & ~(size_t)255;
				size_t a2 = std::max(arenaR, need);
				if (k1 > k0 && a2 * (k1 - k0 + 1) > ctx->mem_budget) break;
				arenaR = a2;
				++k1;
			}
			const int nR = (int)(k1 - k0);
			std::vector<int> sub(redo.begin() + (long)k0, redo.begin() + (long)k1);
			const int blocksR = (nR + groupsPerBlock - 1) / groupsPerBlock;
			if (!db.dp_arena.ensure((size_t)blocksR * groupsPerBlock * arenaR + 16) || !upload(ctx, db.dp_order, sub.data(), sub.size(), strm))
				return ctx->fail("out of device memory (DP program arenas, second attempt)");
			CK(cudaMemsetAsync(cnt, 0, 32, strm));
			CK(cudaMemsetAsync(db.dp_ovf.p, 0, 64, strm));
			KLAUNCH_ON(strm, 10, k_dp2_units, blocksR, 128, ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>(), nR, (char *)db.dp_arena.p, arenaR,
			           (size_t)0, (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt, db.dp_ovf.as<int>(), ctx->coop, (long long *)nullptr, 1);
			++launches;
			CK(cudaGetLastError());
			std::vector<int> hn((size_t)pd.units.size());
			CK(cudaMemcpyAsync(hn.data(), db.dp_ntx.p, sizeof(int) * hn.size(), cudaMemcpyDeviceToHost, strm));
			CK(cudaStreamSynchronize(strm));
			for (int u : sub)
				if (hn[(size_t)u] < 0) still.push_back(u);
			k0 = k1;
		}
		redo.swap(still);
		return 0;
	};
	auto dp_finish = [&](DpPending &pd, cudaStream_t strm, DpBufs &db, int *cnt) -> int {
		if (!pd.active) return 0;
		pd.active = false;
		const int nUnits = (int)pd.units.size();
		int hOvf = 0;
		CK(cudaMemcpyAsync(&hOvf, db.dp_ovf.p, sizeof(int), cudaMemcpyDeviceToHost, strm));
		CK(cudaStreamSynchronize(strm));
		lap("dp program kernel");
		if (dpDbg && useV2) dp_print(pd.units, strm, "program");
		if (hOvf > 0) {
			std::vector<int> hn((size_t)nUnits);
			CK(cudaMemcpy(hn.data(), db.dp_ntx.p, sizeof(int) * (size_t)nUnits, cudaMemcpyDeviceToHost));
			std::vector<int> redo;
			for (int u = 0; u < nUnits; ++u)
				if (hn[(size_t)u] < 0) redo.push_back(u);
			if (dpDbg) {
				CK(cudaMemsetAsync(st.dp_dbg.p, 0, sizeof(long long) * 8 * (size_t)(nUnits + 1), strm));
			}
			if (useV2 && !getenv("PSC_DP_NO_RETRY")) {
				if (dp_program_retry(pd, redo, strm, db, cnt) != 0) return 1;
				lap("dp program kernel, roomy second attempt");
			}
			if (!redo.empty() && dp_walk_units(pd, redo, strm, db, cnt) != 0) return 1;
			lap("dp walk kernel (overflowed units)");
			if (dpDbg) dp_print(pd.units, strm, "walk");
		}
		if (!db.dp_merge_scr.ensure((size_t)pd.mAcc + 16) || !upload(ctx, db.list, pd.comps.data(), pd.comps.size(), strm) ||
		    !upload(ctx, db.dp_scr_off, pd.mergeOff.data(), pd.mergeOff.size(), strm))
			return ctx->fail("out of device memory (DP merge)");
		CK(cudaMemsetAsync(cnt, 0, 32, strm));
#ifndef PSC_CPU_SHIM
		{
			// coop 32: one warp (= one component at a time) per block; coop 1: 128 one-thread groups per block
			const int mThreads = ctx->coop == 32 ? 32 : 128;
			const int mBlocks = ctx->coop == 32 ? std::max(1, std::min(pd.nComp, ctx->n_sm * 12)) : launch_grid(ctx, pd.nComp);
			std::lock_guard<std::mutex> lk_(ctx->launchMu);
			ctx->ktime_begin(7, strm);
			k_dp_merge<<<mBlocks, mThreads, 0, strm>>>(ctx->gd, sv, db.dp_units.as<DpUnit>(), db.list.as<int>(), pd.nComp,
			                                                                 db.dp_ntx.as<int>(), (const char *)db.dp_tx.p, (char *)db.dp_merge_scr.p,
			                                                                 db.dp_scr_off.as<long long>(), cnt, ctx->coop);
			ctx->ktime_end();
		}
#else
		KLAUNCH_ON(strm, 7, k_dp_merge, launch_grid(ctx, pd.nComp), 128, ctx->gd, sv, db.dp_units.as<DpUnit>(), db.list.as<int>(), pd.nComp,
		           db.dp_ntx.as<int>(), (const char *)db.dp_tx.p, (char *)db.dp_merge_scr.p, db.dp_scr_off.as<long long>(), cnt, ctx->coop);
#endif
		++launches;
		CK(cudaGetLastError());
		return 0;
	};
	// issues the DP of `comps` chunk by chunk; all chunks but the last are finished here, the last one is left in `pd`
	auto dp_launch = [&](const std::vector<int> &comps, cudaStream_t strm, DpBufs &db, int *cnt, DpPending &pd, bool heavyLaunch) -> int {
		size_t i0 = 0;
		while (i0 < comps.size()) {
			if (pd.active) {
				if (dp_finish(pd, strm, db, cnt) != 0) return 1;
				CK(cudaStreamSynchronize(strm)); // the merge reads the buffers the next chunk overwrites
			}
			pd.units.clear();
			pd.mergeOff.clear();
			std::vector<long long> cost;
			long long tAcc = 0, mAcc = 0;
			size_t arenaBytes = 0, hotWanted = 0;
			std::vector<size_t> smallNeed; // per unit: arena bytes when run one thread per unit (0: not eligible)
			size_t i1 = i0;
			while (i1 < comps.size()) {
				int c = comps[i1];
				int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
				int W = (n + 63) / 64;
				long long cap = rowOff[(size_t)c + 1] - rowOff[(size_t)c];
				// the reference visits the samples in descending number of constraints (qsort ascending,
				// stable, walked from the back: Q11)
				std::vector<int> ord((size_t)S);
				for (int s = 0; s < S; ++s) ord[(size_t)s] = s;
				std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
					return ctx->samples[(size_t)a].h_ncon[(size_t)c] < ctx->samples[(size_t)b].h_ncon[(size_t)c];
				});
				std::vector<DpUnit> mine;
				long long tNeed = 0;
				size_t aNeed = 0;
				int downsampleCnt = -1;
				for (int oi = S - 1; oi >= 0; --oi) {
					int s = ord[(size_t)oi];
					int C = ctx->samples[(size_t)s].h_ncon[(size_t)c], M = ctx->samples[(size_t)s].h_npair[(size_t)c];
					DpUnit u;
					u.comp = c;
					u.sample = s;
					u.iterBound = oi == S - 1 ? C : 100;
					u.stride = 0;
					if (ctx->prm.max_dp_constraint_size <= 0) {
						bool down = (C > 1000 && (long long)C * 10 < M) || (downsampleCnt > 0 && C >= downsampleCnt) || n >= 1500;
						if (down) { // TranscriptDecider.cpp:3214-3227
							int stride = C > 0 ? M / C : 1;
							if (downsampleCnt > 0) stride = C / downsampleCnt;
							if (stride < 1) stride = 1;
							u.stride = stride;
							if (downsampleCnt <= 0) downsampleCnt = (C + stride - 1) / stride;
						}
					}
					if (u.iterBound < 1) u.iterBound = 1;
					u.scratch_off = 0;
					u.tx_off = tAcc + tNeed;
					tNeed += (long long)dp_txlist_layout(nullptr, nullptr, u.iterBound, W);
					aNeed = std::max(aNeed, dp2_layout(nullptr, nullptr, nullptr, 0, nullptr, 0, n, ctx->h_next_cnt[(size_t)c], std::max(C, 100)));
					hotWanted = std::max(hotWanted, dp2_smem_wanted(n));
					mine.push_back(u);
				}
				long long mNeed = (long long)dp_merge_layout(nullptr, nullptr, W, cap);
				if (i1 > i0 && (size_t)(tAcc + tNeed + mAcc + mNeed) > ctx->mem_budget) break;
				for (auto &u : mine) {
					const int Cu = ctx->samples[(size_t)u.sample].h_ncon[(size_t)c];
					smallNeed.push_back(n <= 64 ? dp2_layout(nullptr, nullptr, nullptr, 0, nullptr, 0, n, ctx->h_next_cnt[(size_t)c], std::max(Cu, 100), 3 * n + 32) : 0);
					// algorithmic bytes of the DP kernel: the sample's constraint table read once, its transcript list written once
					dpBytes += (long long)ctx->samples[(size_t)u.sample].h_ncon[(size_t)c] * (8LL * W + 40) + (long long)u.iterBound * (8LL * W + 8);
					pd.units.push_back(u);
					cost.push_back((long long)u.iterBound * (ctx->samples[(size_t)u.sample].h_ncon[(size_t)c] + 1));
				}
				tAcc += tNeed;
				pd.mergeOff.push_back(mAcc);
				mAcc += mNeed;
				arenaBytes = std::max(arenaBytes, aNeed);
				++i1;
			}
			const int nUnits = (int)pd.units.size();
			pd.nComp = (int)(i1 - i0);
			pd.comps.assign(comps.begin() + (long)i0, comps.begin() + (long)i1);
			pd.mAcc = mAcc;
			std::vector<int> order((size_t)nUnits);
			for (int u = 0; u < nUnits; ++u) order[(size_t)u] = u;
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
			// one arena per cooperative group of the grid; fewer groups when the arenas would not fit the budget.
			// Shared memory: four units per block with a 16 KB slice each (PSC_DP_SMEM_KB overrides).  The launch over the
			// heavy components additionally hands its nCrit most expensive units to the critical stream: one unit per
			// block, with room for the biggest program, on the SMs reserved for them when the device is partitioned.
			arenaBytes = (arenaBytes + 255) & ~(size_t)255;
			const int threads = 128;
			size_t hotBytes = 0, hotCrit = 0;
			int nCrit = 0;
			bool critOnThird = false;
			if (ctx->coop == 32) {
				hotBytes = std::min<size_t>((hotWanted + 1023) & ~(size_t)1023, 16 << 10);
				if (const char *e = getenv("PSC_DP_SMEM_KB")) hotBytes = (size_t)std::max(0, atoi(e)) << 10;
				if (heavyLaunch && useV2 && !dpDbg) {
					nCrit = 64;
					if (const char *e = getenv("PSC_DP_CRIT")) nCrit = std::max(0, atoi(e));
					nCrit = std::min(nCrit, nUnits / 2);
					if (strm == ctx->sCrit) nCrit = nUnits; // the launch already lives in the critical partition: every unit takes that shape
				} else if (!heavyLaunch && useV2 && !dpDbg && nUnits >= 4096) {
					// the bulk launch's own tail (PSC_DP_CRIT_BULK=<n units>, off by default): its most expensive units can also
					// run one per block with a big slice, next to two (instead of three) regular blocks per SM.  Measured on
					// B200: the launch gets slower (321 vs 268 ms), the lost third of the regular blocks costs more.
					nCrit = 0;
					if (const char *e = getenv("PSC_DP_CRIT_BULK")) nCrit = std::max(0, atoi(e));
					nCrit = std::min(nCrit, nUnits / 8);
					critOnThird = nCrit > 0;
				}
				if (nCrit > 0) {
					for (int k = 0; k < nCrit; ++k) {
						int c = pd.units[(size_t)order[(size_t)k]].comp;
						hotCrit = std::max(hotCrit, dp2_smem_wanted(ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c]));
					}
					hotCrit = std::min<size_t>((hotCrit + 1023) & ~(size_t)1023, critOnThird ? (92 << 10) : (110 << 10));
				}
			}
			// The bulk launch hands the light end of its list to a one-thread-per-unit launch: the interpreter is scalar per
			// unit (the lanes of a warp only cooperate in the list building and the assignment scans), so for throughput
			// 32 units per warp beat one; the heavy end stays one unit per warp, where latency counts (PSC_DP_T1_FRAC).
			int nT1 = 0;
			size_t arenaT1 = 0;
			if (ctx->coop == 32 && !heavyLaunch && useV2 && !dpDbg && nUnits >= 4096) {
				double frac = 0.0; // measured on B200: one thread per unit is slower (divergence, uncoalesced tables); kept as a switch
				if (const char *e = getenv("PSC_DP_T1_FRAC")) frac = atof(e);
				const int firstT1 = nUnits - (int)(frac * nUnits);
				std::vector<int> keepW, goT1;
				for (int k = 0; k < nUnits; ++k) {
					int u = order[(size_t)k];
					if (k >= firstT1 && smallNeed[(size_t)u] > 0) {
						goT1.push_back(u);
						arenaT1 = std::max(arenaT1, smallNeed[(size_t)u]);
					} else
						keepW.push_back(u);
				}
				nT1 = (int)goT1.size();
				order = keepW;
				order.insert(order.end(), goT1.begin(), goT1.end());
				arenaT1 = (arenaT1 + 255) & ~(size_t)255;
			}
			const long long groupsPerBlock = threads / ctx->coop;
			const int nBulk = nUnits - nCrit - nT1;
			// packets of the bulk part (PSC_DP_PACKETS=1; default: fixed 16 KB slices)
			std::vector<DpPacket> packets;
			const size_t poolBytes = 72 << 10; // three blocks per SM
			const bool usePackets = ctx->coop == 32 && useV2 && nBulk > 0 && getenv("PSC_DP_PACKETS") && atoi(getenv("PSC_DP_PACKETS")) != 0; // off: measured slower (799 vs 722 ms per step: the block waits for its slowest unit)
			if (usePackets) {
				DpPacket cur;
				int cnt = 0;
				size_t used = 0;
				auto flush = [&]() {
					if (cnt == 0) return;
					for (int q = cnt; q < 4; ++q) { cur.unit[q] = -1; cur.hotOff[q] = 0; cur.hotBytes[q] = 0; }
					packets.push_back(cur);
					cnt = 0;
					used = 0;
				};
				for (int k = nCrit; k < nCrit + nBulk; ++k) {
					const int u = order[(size_t)k];
					const int c = pd.units[(size_t)u].comp;
					size_t need = dp2_smem_slice(ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c]);
					if (need > poolBytes) need = 0; // runs in its global arena
					if (cnt == 4 || used + need > poolBytes) flush();
					cur.unit[cnt] = u;
					cur.hotOff[cnt] = (int)used;
					cur.hotBytes[cnt] = (int)need;
					used += need;
					++cnt;
				}
				flush();
			}
			long long blocks = (nBulk + groupsPerBlock - 1) / groupsPerBlock;
			{
				long long blocksPerSm = critOnThird ? 2 : PSC_DP2_MINBLOCKS;
				if (hotBytes > 0) blocksPerSm = std::max<long long>(1, std::min<long long>(blocksPerSm, (long long)((227 << 10) / (hotBytes * 4 + 1024))));
				blocks = std::min<long long>(blocks, (long long)ctx->n_sm * blocksPerSm);
				long long fit = (long long)(ctx->mem_budget / (arenaBytes * (size_t)groupsPerBlock));
				blocks = std::max<long long>(1, std::min(blocks, fit));
			}
			const size_t bulkArenas = (size_t)blocks * (size_t)groupsPerBlock;
			long long blocksT1 = 0;
			if (nT1 > 0) {
				blocksT1 = std::min<long long>((nT1 + 127) / 128, (long long)ctx->n_sm * PSC_DP2_MINBLOCKS);
				blocksT1 = std::max<long long>(1, std::min<long long>(blocksT1, (long long)(ctx->mem_budget / (arenaT1 * 128))));
			}
			const size_t t1Off = (bulkArenas + (size_t)nCrit) * arenaBytes;
			if (!db.dp_tx.ensure((size_t)tAcc + 16) || !db.dp_ntx.ensure(sizeof(int) * (size_t)(nUnits + 1)) || !db.dp_ovf.ensure(64) ||
			    !db.dp_arena.ensure(t1Off + (size_t)blocksT1 * 128 * arenaT1 + 16))
				return ctx->fail("out of device memory (DP program arenas)");
			CK(cudaMemsetAsync(db.dp_ntx.p, useV2 ? 0 : 0xff, sizeof(int) * (size_t)(nUnits + 1), strm));
			CK(cudaMemsetAsync(db.dp_ovf.p, 0, 64, strm));
			if (!upload(ctx, db.dp_units, pd.units.data(), pd.units.size(), strm) || !upload(ctx, db.dp_order, order.data(), order.size(), strm) ||
			    (usePackets && !upload(ctx, db.dp_packets, packets.data(), packets.size(), strm)))
				return ctx->fail("out of device memory");
			lap("dp host prep");
			CK(cudaMemsetAsync(cnt, 0, 32, strm));
			long long *dbgPtr = nullptr;
			if (dpDbg) {
				if (!st.dp_dbg.ensure(sizeof(long long) * 8 * (size_t)(nUnits + 1))) return ctx->fail("out of device memory");
				CK(cudaMemsetAsync(st.dp_dbg.p, 0, sizeof(long long) * 8 * (size_t)(nUnits + 1), strm));
				dbgPtr = st.dp_dbg.as<long long>();
			}
			if (useV2) {
#ifndef PSC_CPU_SHIM
				CK(cudaFuncSetAttribute(k_dp2_units, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 << 10));
				std::lock_guard<std::mutex> lk_(ctx->launchMu);
				if (nCrit > 0) {
					// the critical units: the setup above must be visible to their stream, and the merge waits for them
					CK(cudaEventRecord(ctx->evCritGo, strm));
					const cudaStream_t sC = critOnThird ? ctx->sThird : ctx->sCrit;
					CK(cudaStreamWaitEvent(sC, ctx->evCritGo, 0));
					ctx->ktime_begin(10, sC);
					k_dp2_units<<<nCrit, 32, hotCrit, sC>>>(ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>(), nCrit,
					                                              (char *)db.dp_arena.p + bulkArenas * arenaBytes, arenaBytes, hotCrit, (char *)db.dp_tx.p,
					                                              db.dp_ntx.as<int>(), cnt + 4, db.dp_ovf.as<int>(), ctx->coop, dbgPtr);
					ctx->ktime_end();
					CK(cudaEventRecord(ctx->evCritDone, sC));
					++launches;
				}
				if (nT1 > 0) {
					CK(cudaEventRecord(ctx->evT1Go, strm));
					CK(cudaStreamWaitEvent(ctx->sThird, ctx->evT1Go, 0));
					ctx->ktime_begin(11, ctx->sThird);
					k_dp2_units<<<(int)blocksT1, 128, 0, ctx->sThird>>>(ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>() + nCrit + nBulk, nT1,
					                                                 (char *)db.dp_arena.p + t1Off, arenaT1, 0, (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt + 5,
					                                                 db.dp_ovf.as<int>(), 1, dbgPtr);
					ctx->ktime_end();
					CK(cudaEventRecord(ctx->evT1Done, ctx->sThird));
					++launches;
				}
				if (usePackets) {
					CK(cudaFuncSetAttribute(k_dp2_packets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poolBytes));
					long long pblocks = std::min<long long>((long long)packets.size(), blocks);
					ctx->ktime_begin(10, strm);
					k_dp2_packets<<<(int)pblocks, 128, poolBytes, strm>>>(ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_packets.as<DpPacket>(), (int)packets.size(),
					                                                   (char *)db.dp_arena.p, arenaBytes, (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt,
					                                                   db.dp_ovf.as<int>(), dbgPtr);
					ctx->ktime_end();
				} else if (nBulk > 0) {
					ctx->ktime_begin(10, strm);
					k_dp2_units<<<(int)blocks, threads, hotBytes * (size_t)(threads / 32), strm>>>(
					    ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>() + nCrit, nBulk, (char *)db.dp_arena.p, arenaBytes, hotBytes,
					    (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt, db.dp_ovf.as<int>(), ctx->coop, dbgPtr);
					ctx->ktime_end();
				}
				if (nCrit > 0) CK(cudaStreamWaitEvent(strm, ctx->evCritDone, 0));
				if (nT1 > 0) CK(cudaStreamWaitEvent(strm, ctx->evT1Done, 0));
#else
				KLAUNCH_ON(strm, 10, k_dp2_units, (int)blocks, threads, ctx->gd, sv, db.dp_units.as<DpUnit>(), db.dp_order.as<int>(), nUnits,
				           (char *)db.dp_arena.p, arenaBytes, hotBytes, (char *)db.dp_tx.p, db.dp_ntx.as<int>(), cnt, db.dp_ovf.as<int>(), ctx->coop, dbgPtr);
#endif
				++launches;
				CK(cudaGetLastError());
			} else {
				int one = nUnits; // every unit goes to the walk engine
				CK(cudaMemcpyAsync(db.dp_ovf.p, &one, sizeof(int), cudaMemcpyHostToDevice, strm));
				CK(cudaStreamSynchronize(strm));
			}
			pd.active = true;
			i0 = i1;
		}
		return 0;
	};
	// The DP kernel ends with a long tail: a handful of (component, sample) units run for most of its duration
	// while the rest of the GPU idles.  So the heaviest DP components (by constraint count) get a launch of their
	// own on the main stream, and everything else -- the DP of the remaining components, its merge, and the
	// per-unit kernels of every other component -- runs next to it on the second stream.
	std::vector<int> heavy, rest;
	{
		const bool forced = getenv("PSC_FORCE_OVERLAP") != nullptr;
		bool split = (ctx->coop == 32 || forced) && !getenv("PSC_NO_OVERLAP") && !getenv("PSC_DEBUG_DP") &&
		             (dpList.size() >= 64 || (forced && dpList.size() >= 2));
		if (split) {
			std::vector<std::pair<int, int>> byC; // (-max constraints over the samples, component)
			for (int c : dpList) {
				int Cmax = 0;
				for (int s = 0; s < S; ++s) Cmax = std::max(Cmax, ctx->samples[(size_t)s].h_ncon[(size_t)c]);
				byC.push_back({-Cmax, c});
			}
			std::sort(byC.begin(), byC.end());
			// the heavy chain runs in the critical SM partition: as many components as keep its (component, sample) units
			// at about twelve per critical SM (PSC_HEAVY_COMPS overrides; measured best on the bench workload: 16 SMs, 20
			// components of 10 samples); without partitioning, 1/16 of the DP components
			size_t nH = std::min<size_t>(std::max<size_t>(dpList.size() / 16, 1), (size_t)ctx->n_sm);
			if (ctx->critSms > 0) nH = std::max<size_t>(1, std::min<size_t>(nH, (size_t)(12 * ctx->critSms / std::max(1, S))));
			if (const char *e = getenv("PSC_HEAVY_COMPS")) nH = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), dpList.size() - 1));
			if (forced && dpList.size() < 64) nH = std::max<size_t>(dpList.size() / 4, 1);
			std::vector<char> isH((size_t)nc, 0);
			for (size_t k = 0; k < nH; ++k) isH[(size_t)byC[k].second] = 1;
			for (int c : dpList) (isH[(size_t)c] ? heavy : rest).push_back(c);
			for (int c : heavy) group[(size_t)c] = 2;
			for (int c : rest) group[(size_t)c] = 1;
		}
	}
	// n_cand of the components of `comps`, copied from the device on `strm` (the two chains own disjoint components)
	auto fetch_n_cand = [&](const std::vector<int> &comps, cudaStream_t strm) -> int {
		std::vector<int> tmp((size_t)nc);
		CK(cudaMemcpyAsync(tmp.data(), sv.n_cand, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, strm));
		CK(cudaStreamSynchronize(strm));
		for (int c : comps) st.h_n_cand[(size_t)c] = tmp[(size_t)c];
		return 0;
	};
	if (!heavy.empty()) {
		// Two chains, each driven by its own host thread so that neither waits for the other's synchronisation points:
		// [main stream]   heavy DP -> merge -> pick / refine / EM / emit of the heavy components
		// [second stream] DP of the rest -> merge -> pick / refine / EM / emit of every other component
		DpPending pdH, pdR;
		// with SM partitioning the whole heavy chain lives in the critical partition
		const cudaStream_t sHeavy = ctx->critSms > 0 ? ctx->sCrit : ctx->sMain;
		if (dp_launch(heavy, sHeavy, st.db[0], counter, pdH, true) != 0) return 1;
		if (dp_launch(rest, ctx->sSecond, st.db[1], counter + 8, pdR, false) != 0) return 1;
		ResultSet &resR = st.passR, &resH = st.passH, &resE = st.passE;
		resR.reset();
		resH.reset();
		resE.reset();
		// A third chain (PSC_THREE_CHAINS=1): the enumerated components need no DP, so their set cover / EM / emit can
		// start right away on a stream of their own.  Off by default -- measured on B200 it does not pay: the DP launch's
		// resident blocks hold the whole register file (3 x 128 threads x 168 registers), so the chain's blocks only
		// get onto the SMs when the DP blocks retire, and they then delay the merge kernels (708 vs 665 ms per step).
		const bool three = getenv("PSC_THREE_CHAINS") && atoi(getenv("PSC_THREE_CHAINS")) != 0;
		int rcR = 0, rcE = 0;
		std::thread second([&]() {
			cudaSetDevice(ctx->device);
			rcR = [&]() -> int {
				if (dp_finish(pdR, ctx->sSecond, st.db[1], counter + 8) != 0) return 1;
				if (fetch_n_cand(rest, ctx->sSecond) != 0) return 1;
				lap("dp merge (second stream)");
				if (run_units(three ? 4 : 1, ctx->sSecond, st.ub[1], counter + 8, resR) != 0) return 1;
				lap("second stream done");
				return 0;
			}();
		});
		std::thread third;
		if (three)
			third = std::thread([&]() {
				cudaSetDevice(ctx->device);
				rcE = run_units(3, ctx->sThird, st.ub[2], counter + 16, resE);
			});
		int rcH = [&]() -> int {
			if (dp_finish(pdH, sHeavy, st.db[0], counter) != 0) return 1;
			if (fetch_n_cand(heavy, sHeavy) != 0) return 1;
			lap("heavy dp merge");
			return run_units(2, sHeavy, st.ub[0], counter, resH);
		}();
		second.join();
		if (third.joinable()) third.join();
		if (rcH != 0 || rcR != 0 || rcE != 0) return 1;
		lap("all chains done");
		if (three) {
			merge_passes(st.passRE, resE, resR);
			merge_passes(st, st.passRE, resH);
		} else
		merge_passes(st, resR, resH);
	} else {
		if (!dpList.empty()) {
			DpPending pd;
			if (dp_launch(dpList, ctx->sMain, st.db[1], counter, pd, false) != 0) return 1;
			if (dp_finish(pd, ctx->sMain, st.db[1], counter) != 0) return 1;
			if (fetch_n_cand(dpList, ctx->sMain) != 0) return 1;
		}
		if (run_units(0, ctx->sMain, st.ub[0], counter, st) != 0) return 1;
	}
	lap("exon offsets");
	int herr = 0;
	CK(cudaMemcpyAsync(&herr, sv.error_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->sMain));
	CK(cudaEventRecord(ctx->ev1, ctx->sMain));
	CK(cudaStreamSynchronize(ctx->sMain));
	if (herr != 0) return ctx->fail("device-side capacity error " + std::to_string(herr));
	float ms = 0;
	cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
	ctx->ktime_collect();
	{
		// algorithmic bytes of the per-unit kernels: the constraint + pair tables are read
		// once, the compatibility bit-matrix is written and read once, the picked transcripts written once
		long long bytes = 0;
		for (int c = 0; c < nc; ++c) {
			int n = ctx->h_comp_off[(size_t)c + 1] - ctx->h_comp_off[(size_t)c];
			int W = (n + 63) / 64;
			long long A = st.h_n_cand[(size_t)c];
			bytes += 48LL * n + 8LL * ctx->h_next_cnt[(size_t)c] + A * (8LL * W + 16);
			for (int s = 0; s < S; ++s) {
				long long C = ctx->samples[(size_t)s].h_ncon[(size_t)c], M = ctx->samples[(size_t)s].h_npair[(size_t)c];
				bytes += C * (8LL * W + 40) + M * 24 + 2 * ((A * M + 7) / 8);
			}
		}
		bytes += (long long)st.r_comp.size() * 40 + (long long)st.r_exon_start.size() * 8;
		ctx->timing.kernel_bytes[5] = bytes;
		ctx->timing.kernel_bytes[8] = bytes;
		ctx->timing.kernel_bytes[4] = dpBytes;
		ctx->timing.kernel_bytes[10] = dpBytes; // the static-program engine reads / writes the same tables
	}
	ctx->timing.solve_ms = ms;
	ctx->timing.kernel_launches += launches;
	ctx->solved = true;
	return 0;
}

int psc_transcript_counts(psc_ctx *ctx, int64_t *n_transcripts, int64_t *n_exons)
{
	if (!ctx->solved) return ctx->fail("psc_solve has not run");
	*n_transcripts = (int64_t)ctx->solve.r_comp.size();
	*n_exons = (int64_t)ctx->solve.r_exon_start.size();
	return 0;
}

int psc_get_transcripts(psc_ctx *ctx, psc_transcripts *out)
{
	if (!ctx->solved) return ctx->fail("psc_solve has not run");
	SolveState &st = ctx->solve;
	size_t nt = st.r_comp.size(), ne = st.r_exon_start.size();
	out->n_transcripts = (int64_t)nt;
	out->n_exons = (int64_t)ne;
	if (nt) {
		memcpy(out->comp, st.r_comp.data(), sizeof(int) * nt);
		memcpy(out->sample, st.r_sample.data(), sizeof(int) * nt);
		memcpy(out->chr, st.r_chr.data(), sizeof(int) * nt);
		memcpy(out->gene_id, st.r_gene.data(), sizeof(int) * nt);
		memcpy(out->strand, st.r_strand.data(), nt);
		memcpy(out->fpkm, st.r_fpkm.data(), sizeof(double) * nt);
		memcpy(out->cov, st.r_cov.data(), sizeof(double) * nt);
	}
	for (size_t k = 0; k <= nt; ++k) out->exon_off[k] = (int64_t)st.r_exon_off[k];
	if (ne) {
		memcpy(out->exon_start, st.r_exon_start.data(), sizeof(int) * ne);
		memcpy(out->exon_end, st.r_exon_end.data(), sizeof(int) * ne);
	}
	return 0;
}

int psc_get_comp_stats(psc_ctx *ctx, psc_comp_stats *out)
{
	if (!ctx->solved) return ctx->fail("psc_solve has not run");
	size_t nc = (size_t)ctx->gd.n_comp;
	if (nc) {
		memcpy(out->at_cnt, ctx->solve.h_at_cnt.data(), sizeof(int) * nc);
		memcpy(out->use_dp, ctx->solve.h_use_dp.data(), sizeof(int) * nc);
		memcpy(out->n_candidates, ctx->solve.h_n_cand.data(), sizeof(int) * nc);
	}
	return 0;
}

} // extern "C"
