// See bam_io.h.  BGZF framing per the SAM/BAM specification; the record filter restates
// Alignments::Next() (/root/reference/alignments.hpp:146-277) and the library statistics restate
// Alignments::GetGeneralInfo (alignments.hpp:489-571).
#include "bam_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace psc {

uint64_t hash_read_id(const char *s, size_t n)
{
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; ++i) {
		h ^= (unsigned char)s[i];
		h *= 1099511628211ull;
	}
	// finaliser (splitmix64) so that near-identical ids spread over all 64 bits
	h += 0x9e3779b97f4a7c15ull;
	h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ull;
	h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
	h ^= h >> 31;
	return h ? h : 1;
}

namespace {

struct Block {
	size_t coff;   // offset of the deflate payload in the file image
	uint32_t clen; // payload bytes
	uint32_t isize;
	uint64_t uoff; // offset in the inflated image
};

// read-only view of a whole file: mapped (no copy; the page cache is the buffer), or read if it cannot be mapped
struct FileView {
	const uint8_t *p = nullptr;
	size_t n = 0;
	bool mapped = false;
	std::vector<uint8_t> owned;
	~FileView()
	{
		if (mapped && p) munmap(const_cast<uint8_t *>(p), n);
	}
	const uint8_t *data() const { return p; }
	size_t size() const { return n; }
	uint8_t operator[](size_t i) const { return p[i]; }
};

bool load_file(const std::string &path, std::vector<uint8_t> &buf, std::string &err);

bool open_view(const std::string &path, FileView &v, std::string &err)
{
	int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) {
		err = "Can not open " + path + ".";
		return false;
	}
	struct stat st;
	if (fstat(fd, &st) == 0 && st.st_size > 0) {
		void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
		if (m != MAP_FAILED) {
			madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
			v.p = (const uint8_t *)m;
			v.n = (size_t)st.st_size;
			v.mapped = true;
			close(fd);
			return true;
		}
	}
	close(fd);
	if (!load_file(path, v.owned, err)) return false;
	v.p = v.owned.data();
	v.n = v.owned.size();
	return true;
}

bool load_file(const std::string &path, std::vector<uint8_t> &buf, std::string &err)
{
	FILE *fp = fopen(path.c_str(), "rb");
	if (!fp) {
		err = "Can not open " + path + ".";
		return false;
	}
	fseek(fp, 0, SEEK_END);
	long sz = ftell(fp);
	fseek(fp, 0, SEEK_SET);
	buf.resize((size_t)sz);
	size_t got = sz ? fread(buf.data(), 1, (size_t)sz, fp) : 0;
	fclose(fp);
	if (got != (size_t)sz) {
		err = "short read on " + path;
		return false;
	}
	return true;
}

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

template <class Bytes> bool index_blocks(const Bytes &f, std::vector<Block> &blocks, uint64_t &total,
                  std::string &err, size_t max_blocks = (size_t)-1)
{
	size_t off = 0;
	total = 0;
	while (off + 18 <= f.size() && blocks.size() < max_blocks) {
		const uint8_t *p = f.data() + off;
		if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) {
			err = "not a BGZF file";
			return false;
		}
		uint16_t xlen = rd16(p + 10);
		size_t x = off + 12, xend = x + xlen;
		int bsize = -1;
		while (x + 4 <= xend && xend <= f.size()) {
			uint16_t slen = rd16(f.data() + x + 2);
			if (f[x] == 'B' && f[x + 1] == 'C' && slen == 2)
				bsize = rd16(f.data() + x + 4);
			x += 4 + slen;
		}
		if (bsize < 0 || off + (size_t)bsize + 1 > f.size()) {
			err = "truncated BGZF block";
			return false;
		}
		Block b;
		b.coff = off + 12 + xlen;
		b.clen = (uint32_t)(bsize + 1 - 12 - xlen - 8);
		b.isize = rd32(f.data() + off + bsize + 1 - 4);
		b.uoff = total;
		total += b.isize;
		blocks.push_back(b);
		off += (size_t)bsize + 1;
	}
	return true;
}

bool inflate_block(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t isize)
{
	if (isize == 0)
		return true;
	z_stream zs;
	memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, -15) != Z_OK)
		return false;
	zs.next_in = const_cast<uint8_t *>(src);
	zs.avail_in = clen;
	zs.next_out = dst;
	zs.avail_out = isize;
	int rc = inflate(&zs, Z_FINISH);
	inflateEnd(&zs);
	return rc == Z_STREAM_END && zs.avail_out == 0;
}

template <class Bytes> bool inflate_all(const Bytes &f, const std::vector<Block> &blocks, uint64_t total,
                 int threads, std::vector<uint8_t> &out, std::string &err)
{
	out.resize(total);
	std::atomic<size_t> next(0);
	std::atomic<bool> ok(true);
	auto work = [&]() {
		for (;;) {
			size_t i = next.fetch_add(64);
			if (i >= blocks.size())
				break;
			size_t e = std::min(blocks.size(), i + 64);
			for (; i < e; ++i)
				if (!inflate_block(f.data() + blocks[i].coff, blocks[i].clen,
				                   out.data() + blocks[i].uoff, blocks[i].isize))
					ok = false;
		}
	};
	if (threads <= 1) {
		work();
	} else {
		std::vector<std::thread> pool;
		for (int t = 0; t < threads; ++t)
			pool.emplace_back(work);
		for (auto &t : pool)
			t.join();
	}
	if (!ok) {
		err = "BGZF inflate failed";
		return false;
	}
	return true;
}

// returns pointer to the value byte (type char) of aux tag `tag`, or NULL
const uint8_t *aux_find(const uint8_t *p, const uint8_t *end, const char tag[2])
{
	while (p + 3 <= end) {
		bool hit = p[0] == (uint8_t)tag[0] && p[1] == (uint8_t)tag[1];
		uint8_t type = p[2];
		const uint8_t *v = p + 3;
		if (hit)
			return p + 2;
		switch (type) {
		case 'A': case 'c': case 'C': v += 1; break;
		case 's': case 'S': v += 2; break;
		case 'i': case 'I': case 'f': v += 4; break;
		case 'd': v += 8; break;
		case 'Z': case 'H':
			while (v < end && *v) ++v;
			++v;
			break;
		case 'B': {
			if (v + 5 > end) return NULL;
			uint8_t st = v[0];
			uint32_t n = rd32(v + 1);
			int w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
			v += 5 + (size_t)n * w;
			break;
		}
		default: return NULL;
		}
		p = v;
	}
	return NULL;
}

// bam_aux2i semantics (samtools-0.1.19/bam_aux.c)
int32_t aux_int(const uint8_t *t)
{
	const uint8_t *v = t + 1;
	switch (*t) {
	case 'c': return (int8_t)v[0];
	case 'C': return v[0];
	case 's': return (int16_t)rd16(v);
	case 'S': return rd16(v);
	case 'i': case 'I': return (int32_t)rd32(v);
	default: return 0;
	}
}

bool parse_header(const uint8_t *u, size_t n, size_t &off, std::vector<std::string> &names,
                  std::vector<int32_t> &lens, std::string &err)
{
	if (n < 12 || memcmp(u, "BAM\1", 4) != 0) {
		err = "bad BAM magic";
		return false;
	}
	uint32_t l_text = rd32(u + 4);
	off = 8 + (size_t)l_text;
	if (off + 4 > n) { err = "truncated BAM header"; return false; }
	uint32_t n_ref = rd32(u + off);
	off += 4;
	for (uint32_t i = 0; i < n_ref; ++i) {
		if (off + 4 > n) { err = "truncated BAM header"; return false; }
		uint32_t l_name = rd32(u + off);
		off += 4;
		if (off + l_name + 4 > n) { err = "truncated BAM header"; return false; }
		names.emplace_back((const char *)(u + off), l_name ? l_name - 1 : 0);
		off += l_name;
		lens.push_back((int32_t)rd32(u + off));
		off += 4;
	}
	return true;
}

} // namespace

bool read_bam_header(const std::string &path, std::vector<std::string> &names,
                     std::vector<int32_t> &lens, std::string &err)
{
	FileView f;
	if (!open_view(path, f, err))
		return false;
	// the header may span several blocks; inflate progressively (only the first blocks are indexed)
	std::vector<Block> blocks;
	uint64_t total = 0;
	if (!index_blocks(f, blocks, total, err, 4096))
		return false;
	std::vector<uint8_t> u;
	size_t used = 0;
	size_t nb = 1;
	for (;;) {
		nb = std::min(nb, blocks.size());
		if (nb == 0)
			break;
		std::vector<Block> part(blocks.begin(), blocks.begin() + nb);
		uint64_t t = part.back().uoff + part.back().isize;
		if (!inflate_all(f, part, t, 1, u, err))
			return false;
		names.clear();
		lens.clear();
		std::string e2;
		if (parse_header(u.data(), u.size(), used, names, lens, e2))
			return true;
		if (nb == blocks.size()) {
			err = e2;
			return false;
		}
		nb *= 4;
	}
	err = "empty BAM";
	return false;
}

// -------------------------------------------------------------------------------------------
// ReadAccumulator: the per-record filter of Alignments::Next() + GetGeneralInfo(true)
// -------------------------------------------------------------------------------------------
struct ReadAccumulator::State {
	std::vector<int32_t> lens, mateDiff;
	int infoCnt = 0;
	bool infoDone = false;
	uint32_t prim = 0;
	int cur_chr = 0;
};

ReadAccumulator::ReadAccumulator(SampleReads &o) : out(o), st(new State)
{
	out.chr_off.assign(out.chr_names.size() + 1, 0);
	out.seg_off.assign(1, 0);
}
ReadAccumulator::~ReadAccumulator() { delete st; }

void ReadAccumulator::add(int32_t tid, int32_t pos, uint32_t flag, int32_t l_seq, int32_t mtid,
                          int32_t mpos, const char *qname, size_t ln, const uint32_t *cig,
                          uint32_t n_cig, bool has_nh, int32_t nh_tag)
{
	const int sampleMax = 1000000;
	int n_chr = (int)out.chr_names.size();
	bool primary = (flag & 0x900) == 0;
	if (primary)
		++st->prim; // alignments.hpp:170-171 (counted before every filter)
	if (flag & 0xC)
		return;
	// ---- GetGeneralInfo(true): first 1M primary, mapped-pair records (alignments.hpp:500-537) ----
	if (!st->infoDone && primary) {
		if ((int)st->lens.size() < sampleMax)
			st->lens.push_back(l_seq);
		if ((int)st->mateDiff.size() < sampleMax && tid == mtid && pos < mpos)
			st->mateDiff.push_back(mpos - pos);
		if (++st->infoCnt >= sampleMax)
			st->infoDone = true;
	}
	// ---- Alignments::Next() filters (alignments.hpp:180-271) ----
	int nh = 1;
	if (has_nh) {
		nh = nh_tag;
		if (nh >= 5)
			return;
	}
	int32_t qlen = 0;
	for (uint32_t i = 0; i < n_cig; ++i) {
		uint32_t op = cig[i] & 0xf;
		if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8)
			qlen += (int32_t)(cig[i] >> 4);
	}
	if (qlen != l_seq)
		return;
	bool clipMiddle = false, clipHead = false, clipTail = false, tooMany = false;
	int segCnt = 0;
	int32_t sa[128], sb[128];
	int32_t start = pos, len = 0;
	for (uint32_t i = 0; i < n_cig && !tooMany; ++i) {
		uint32_t op = cig[i] & 0xf;
		int32_t num = (int32_t)(cig[i] >> 4);
		switch (op) {
		case 0: case 2: len += num; break;
		case 4: case 5: case 6:
			if (i == 0) clipHead = true;
			else if (i == n_cig - 1) clipTail = true;
			else clipMiddle = true;
			break;
		case 1: break;
		case 3:
			if (segCnt >= 127) { tooMany = true; break; }
			sa[segCnt] = start;
			sb[segCnt] = start + len - 1;
			++segCnt;
			start = start + len + num;
			len = 0;
			break;
		default: len += num; break;
		}
	}
	if (clipMiddle || tooMany)
		return;
	if (len > 0) {
		if (segCnt >= 127)
			return; // the reference would overrun segments[MAX_SEG_COUNT]; such reads are dropped
		sa[segCnt] = start;
		sb[segCnt] = start + len - 1;
		++segCnt;
	}
	if (segCnt == 0)
		return; // nothing aligned (the reference would reuse stale segments); dropped
	if (mtid == tid) { // mate inside one of the read's own introns: alignments.hpp:262-271
		int i;
		for (i = 0; i < segCnt - 1; ++i)
			if (mpos >= sb[i] && mpos <= sa[i + 1])
				break;
		if (i < segCnt - 1)
			return;
	}
	if (tid < 0 || tid >= n_chr)
		return;
	// ---- emit ----
	while (st->cur_chr < tid) {
		++st->cur_chr;
		out.chr_off[st->cur_chr] = (int64_t)out.pos.size();
	}
	uint8_t fl = 0;
	if (primary) fl |= RF_PRIMARY;
	if (clipHead) fl |= RF_CLIP_HEAD;
	if (clipTail) fl |= RF_CLIP_TAIL;
	if (flag & 0x40) fl |= RF_FIRST_MATE;
	if (!(nh > 1)) fl |= RF_UNIQUE;
	bool mateSame = (mtid == tid); // flag&0x8 already excluded above
	if (mateSame) fl |= RF_MATE_SAME_CHR;
	uint64_t key = 0;
	if (mateSame && mpos != sa[0]) {
		if (ln > 254) ln = 254;
		if (mpos > sa[0]) {
			key = hash_read_id(qname, ln); // MateReadIds::Insert stores the id verbatim
		} else {
			// MateReadIds::Query flips a trailing [./][12] (Constraints.hpp:196-205; always on, Q2)
			char tmp[256];
			memcpy(tmp, qname, ln);
			if (ln >= 2 && (tmp[ln - 1] == '1' || tmp[ln - 1] == '2') &&
			    (tmp[ln - 2] == '.' || tmp[ln - 2] == '/'))
				tmp[ln - 1] = (char)('2' - tmp[ln - 1] + '1');
			key = hash_read_id(tmp, ln);
		}
	}
	out.pos.push_back(sa[0]);
	out.mate_pos.push_back(mpos);
	out.flags.push_back(fl);
	out.nh.push_back((uint8_t)nh);
	out.name_key.push_back(key);
	out.prim_cum.push_back(st->prim);
	for (int i = 0; i < segCnt; ++i) {
		out.seg_a.push_back(sa[i]);
		out.seg_b.push_back(sb[i]);
	}
	out.seg_off.push_back((uint32_t)out.seg_a.size());
}

void ReadAccumulator::finish()
{
	int n_chr = (int)out.chr_names.size();
	while (st->cur_chr < n_chr) {
		++st->cur_chr;
		out.chr_off[st->cur_chr] = (int64_t)out.pos.size();
	}
	out.total_primary = st->prim;
	// ---- library statistics (alignments.hpp:539-567) ----
	std::vector<int32_t> &lens = st->lens, &mateDiff = st->mateDiff;
	if (!lens.empty())
		out.read_len = *std::max_element(lens.begin(), lens.end());
	else
		out.read_len = 150;
	if (!mateDiff.empty()) {
		out.mate_paired = 1;
		std::sort(mateDiff.begin(), mateDiff.end());
		long long sum = 0, sumsq = 0;
		int i;
		int cnt = (int)mateDiff.size();
		for (i = 0; i < cnt * 0.7; ++i) {
			int32_t v = mateDiff[i] + out.read_len;
			sum += v;
			sumsq += (int32_t)((uint32_t)v * (uint32_t)v); // int*int in the reference (wraps)
		}
		int k = i;
		out.frag_len = (int)(sum / k);
		long long var = sumsq / k - (long long)(int32_t)((uint32_t)out.frag_len * (uint32_t)out.frag_len);
		out.frag_stdev = (int)std::sqrt((double)var);
	} else {
		out.mate_paired = 0;
		out.frag_len = out.read_len;
		out.frag_stdev = 0;
	}
}

// Single pass over one BAM: the BGZF blocks are inflated by `threads` workers, in order, into one image while this
// thread already parses the records behind them (alignments.hpp:146-277 and :489-571 need one sequential reader; the
// inflate, which is most of the time, does not).
bool decode_bam(const std::string &path, int threads, SampleReads &out, std::string &err)
{
	FileView f;
	if (!open_view(path, f, err))
		return false;
	std::vector<Block> blocks;
	uint64_t total = 0;
	if (!index_blocks(f, blocks, total, err))
		return false;
	// The inflated image is an anonymous mapping: pages the parser is done with are handed back (MADV_DONTNEED) and the
	// inflate workers stay at most WINDOW chunks ahead of it, so a sample holds ~0.4 GB of inflated data at a time
	// instead of all of it (2.8 GB per sample on the bench workload).
	struct Mapping {
		uint8_t *p = nullptr;
		size_t n = 0;
		~Mapping() { if (p) munmap(p, n); }
		uint8_t *get() const { return p; }
	} image;
	image.n = (size_t)total + 4096;
	{
		void *m = mmap(nullptr, image.n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (m == MAP_FAILED) {
			err = "out of memory for the inflated BAM image";
			return false;
		}
		image.p = (uint8_t *)m;
	}
	const size_t CHUNK = 64;   // blocks per unit of work (~4 MB inflated)
	const size_t WINDOW = 96;  // chunks the inflate may run ahead of the parser
	std::atomic<size_t> consumedChunk(0);
	std::atomic<bool> stop(false);
	const size_t nChunks = (blocks.size() + CHUNK - 1) / CHUNK;
	std::unique_ptr<std::atomic<unsigned char>[]> chunkDone(new std::atomic<unsigned char>[nChunks + 1]);
	for (size_t k = 0; k <= nChunks; ++k) chunkDone[k].store(0, std::memory_order_relaxed);
	std::atomic<size_t> next(0);
	std::atomic<bool> ok(true);
	auto work = [&]() {
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		bool init = inflateInit2(&zs, -15) == Z_OK;
		for (;;) {
			size_t k = next.fetch_add(1);
			if (k >= nChunks) break;
			while (k >= consumedChunk.load(std::memory_order_acquire) + WINDOW && !stop.load(std::memory_order_relaxed))
				std::this_thread::sleep_for(std::chrono::microseconds(50));
			if (stop.load(std::memory_order_relaxed)) break;
			size_t i = k * CHUNK, e = std::min(blocks.size(), i + CHUNK);
			for (; i < e && init; ++i) {
				const Block &b = blocks[i];
				if (b.isize == 0) continue;
				inflateReset(&zs);
				zs.next_in = const_cast<uint8_t *>(f.data() + b.coff);
				zs.avail_in = b.clen;
				zs.next_out = image.get() + b.uoff;
				zs.avail_out = b.isize;
				int rc = inflate(&zs, Z_FINISH);
				if (!(rc == Z_STREAM_END && zs.avail_out == 0)) ok = false;
			}
			if (!init) ok = false;
			chunkDone[k].store(1, std::memory_order_release);
		}
		if (init) inflateEnd(&zs);
	};
	std::vector<std::thread> pool;
	for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(work);
	struct Joiner {
		std::vector<std::thread> &p;
		std::atomic<bool> &stop;
		~Joiner()
		{
			stop.store(true); // an early return must not leave workers waiting for the parser
			for (auto &t : p) if (t.joinable()) t.join();
		}
	} joiner{pool, stop};
	// chunks [0, relChunk) lie entirely behind the parser; their pages up to releasedUpTo have been handed back
	size_t relChunk = 0, releasedUpTo = 0;
	auto release_behind = [&](size_t off) {
		while (relChunk < nChunks) {
			size_t b = std::min(blocks.size(), (relChunk + 1) * CHUNK);
			uint64_t chunkEnd = b < blocks.size() ? blocks[b].uoff : total;
			if (chunkEnd > off) break;
			++relChunk;
			size_t lim = (size_t)chunkEnd & ~(size_t)4095;
			if (lim > releasedUpTo) {
				madvise(image.get() + releasedUpTo, lim - releasedUpTo, MADV_DONTNEED);
				releasedUpTo = lim;
			}
			consumedChunk.store(relChunk, std::memory_order_release);
		}
	};
	// bytes [0, ready) of the image are final
	size_t frontChunk = 0;
	uint64_t ready = 0;
	auto wait_for = [&](uint64_t need) -> bool {
		while (ready < need) {
			if (frontChunk >= nChunks) return false;
			if (chunkDone[frontChunk].load(std::memory_order_acquire)) {
				++frontChunk;
				size_t b = std::min(blocks.size(), frontChunk * CHUNK);
				ready = b < blocks.size() ? blocks[b].uoff : total;
			} else
				std::this_thread::sleep_for(std::chrono::microseconds(50));
		}
		return true;
	};
	const uint8_t *ubase = image.get();
	const size_t usize = (size_t)total;

	out = SampleReads();
	out.path = path;
	size_t off = 0;
	{
		// the header: magic, l_text, text, n_ref, references -- wait for as much as it needs
		if (!wait_for(std::min<uint64_t>(total, 12))) { err = "truncated BAM header"; return false; }
		uint64_t need = 12;
		if (usize >= 8) need = 12 + (uint64_t)rd32(ubase + 4);
		if (!wait_for(std::min<uint64_t>(total, need))) { err = "truncated BAM header"; return false; }
		// the reference list follows the text; its size is only known entry by entry: wait for the whole first chunks
		// until the parse succeeds
		for (;;) {
			size_t o2 = 0;
			std::vector<std::string> names;
			std::vector<int32_t> lens;
			std::string e2;
			if (parse_header(ubase, (size_t)ready, o2, names, lens, e2)) {
				off = o2;
				out.chr_names = names;
				out.chr_len = lens;
				break;
			}
			if (ready >= total) { err = e2; return false; }
			if (!wait_for(std::min<uint64_t>(total, ready + 1))) { err = e2; return false; }
		}
	}
	ReadAccumulator acc(out);
	const uint8_t *base = ubase;
	size_t n = usize;
	std::vector<uint32_t> cigbuf;
	while (off + 4 <= n) {
		release_behind(off);
		if (!wait_for(off + 4)) break;
		uint32_t bs = rd32(base + off);
		if (off + 4 + (size_t)bs <= n && !wait_for(off + 4 + (size_t)bs)) break;
		if (off + 4 + bs > n || bs < 32) {
			err = "truncated BAM record";
			return false;
		}
		const uint8_t *r = base + off + 4;
		off += 4 + (size_t)bs;
		int32_t tid = (int32_t)rd32(r);
		int32_t pos = (int32_t)rd32(r + 4);
		uint32_t l_name = r[8];
		uint32_t n_cig = rd16(r + 12);
		uint32_t flag = rd16(r + 14);
		int32_t l_seq = (int32_t)rd32(r + 16);
		int32_t mtid = (int32_t)rd32(r + 20);
		int32_t mpos = (int32_t)rd32(r + 24);
		const char *qname = (const char *)(r + 32);
		const uint8_t *cig = r + 32 + l_name;
		const uint8_t *aux = cig + 4 * (size_t)n_cig + (size_t)((l_seq + 1) / 2) + (size_t)l_seq;
		const uint8_t *rend = r + bs;
		cigbuf.resize(n_cig);
		for (uint32_t i = 0; i < n_cig; ++i)
			cigbuf[i] = rd32(cig + 4 * i);
		const uint8_t *t = aux < rend ? aux_find(aux, rend, "NH") : NULL;
		acc.add(tid, pos, flag, l_seq, mtid, mpos, qname, strnlen(qname, l_name), cigbuf.data(), n_cig,
		        t != NULL, t ? aux_int(t) : 1);
	}
	acc.finish();
	consumedChunk.store(nChunks, std::memory_order_release); // a trailing block without records: let the workers finish
	for (auto &t : pool) t.join();
	if (!ok) {
		err = "BGZF inflate failed";
		return false;
	}
	return true;
}

// -------------------------------------------------------------------------------------------
// writer
// -------------------------------------------------------------------------------------------
namespace {
struct WriterImpl {
	FILE *fp = nullptr;
	std::vector<uint8_t> buf;
	void flush_block()
	{
		size_t done = 0;
		while (done < buf.size()) {
			size_t chunk = std::min<size_t>(buf.size() - done, 0xff00);
			uint8_t outb[0x10000 + 64];
			z_stream zs;
			memset(&zs, 0, sizeof(zs));
			deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
			zs.next_in = buf.data() + done;
			zs.avail_in = (uInt)chunk;
			zs.next_out = outb + 18;
			zs.avail_out = sizeof(outb) - 18 - 8;
			deflate(&zs, Z_FINISH);
			size_t clen = zs.total_out;
			deflateEnd(&zs);
			static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
			memcpy(outb, hdr, 16);
			size_t bsize = 18 + clen + 8;
			outb[16] = (uint8_t)((bsize - 1) & 0xff);
			outb[17] = (uint8_t)((bsize - 1) >> 8);
			uint32_t crc = (uint32_t)crc32(crc32(0, NULL, 0), buf.data() + done, (uInt)chunk);
			uint8_t *tail = outb + 18 + clen;
			for (int i = 0; i < 4; ++i) tail[i] = (uint8_t)(crc >> (8 * i));
			for (int i = 0; i < 4; ++i) tail[4 + i] = (uint8_t)((uint32_t)chunk >> (8 * i));
			fwrite(outb, 1, bsize, fp);
			done += chunk;
		}
		buf.clear();
	}
	void put(const void *p, size_t n)
	{
		const uint8_t *b = (const uint8_t *)p;
		buf.insert(buf.end(), b, b + n);
		if (buf.size() >= 0xff00 * 8)
			flush_block();
	}
	void put32(uint32_t v) { put(&v, 4); }
};
} // namespace

bool BamWriter::open(const std::string &path, const std::vector<std::string> &names,
                     const std::vector<int32_t> &lens)
{
	WriterImpl *w = new WriterImpl;
	w->fp = fopen(path.c_str(), "wb");
	if (!w->fp) {
		delete w;
		return false;
	}
	impl = w;
	std::string text = "@HD\tVN:1.0\tSO:coordinate\n";
	for (size_t i = 0; i < names.size(); ++i)
		text += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
	w->put("BAM\1", 4);
	w->put32((uint32_t)text.size());
	w->put(text.data(), text.size());
	w->put32((uint32_t)names.size());
	for (size_t i = 0; i < names.size(); ++i) {
		w->put32((uint32_t)names[i].size() + 1);
		w->put(names[i].c_str(), names[i].size() + 1);
		w->put32((uint32_t)lens[i]);
	}
	w->flush_block();
	return true;
}

static int reg2bin(int beg, int end)
{
	--end;
	if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
	if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
	if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
	if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
	if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
	return 0;
}

void BamWriter::write(int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, const std::string &name,
                      const uint32_t *cigar, int n_cigar, int32_t l_seq, int32_t mtid, int32_t mpos,
                      int32_t tlen, const std::string &aux)
{
	WriterImpl *w = (WriterImpl *)impl;
	int32_t reflen = 0;
	for (int i = 0; i < n_cigar; ++i) {
		uint32_t op = cigar[i] & 0xf;
		if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8)
			reflen += (int32_t)(cigar[i] >> 4);
	}
	uint32_t l_name = (uint32_t)name.size() + 1;
	uint32_t bs = 32 + l_name + 4 * (uint32_t)n_cigar + (uint32_t)((l_seq + 1) / 2) + (uint32_t)l_seq +
	              (uint32_t)aux.size();
	std::vector<uint8_t> rec(4 + bs);
	uint8_t *p = rec.data();
	auto w32 = [&](size_t o, uint32_t v) { memcpy(p + o, &v, 4); };
	w32(0, bs);
	w32(4, (uint32_t)tid);
	w32(8, (uint32_t)pos);
	uint32_t bin = (uint32_t)reg2bin(pos, pos + (reflen > 0 ? reflen : 1));
	w32(12, (bin << 16) | ((uint32_t)mapq << 8) | l_name);
	w32(16, ((uint32_t)flag << 16) | (uint32_t)n_cigar);
	w32(20, (uint32_t)l_seq);
	w32(24, (uint32_t)mtid);
	w32(28, (uint32_t)mpos);
	w32(32, (uint32_t)tlen);
	size_t o = 36;
	memcpy(p + o, name.c_str(), l_name);
	o += l_name;
	memcpy(p + o, cigar, 4 * (size_t)n_cigar);
	o += 4 * (size_t)n_cigar;
	memset(p + o, 0x11, (size_t)((l_seq + 1) / 2)); // 'A' = 1
	o += (size_t)((l_seq + 1) / 2);
	memset(p + o, 30, (size_t)l_seq);
	o += (size_t)l_seq;
	if (!aux.empty())
		memcpy(p + o, aux.data(), aux.size());
	w->put(rec.data(), rec.size());
}

void BamWriter::close()
{
	WriterImpl *w = (WriterImpl *)impl;
	if (!w)
		return;
	if (!w->buf.empty())
		w->flush_block();
	static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
	                                0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	fwrite(eof, 1, 28, w->fp);
	fclose(w->fp);
	delete w;
	impl = nullptr;
}

} // namespace psc
