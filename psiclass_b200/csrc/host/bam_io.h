// BAM/BGZF reader + writer and the alignment-record decoder of the stage-3 host side.
//
// Replaces, for the stage-3 path only, the reference's `Alignments` class
// (/root/reference/alignments.hpp:16-572) and the slice of samtools-0.1.19 it uses
// (samopen/samread/bam_aux_get).  The decoder performs ONE pass per BAM and produces
//   * the library statistics of Alignments::GetGeneralInfo(true) (alignments.hpp:489-571)
//   * the compact, structure-of-arrays alignment records the CUDA kernels consume, holding
//     exactly the reads for which Alignments::Next() (alignments.hpp:146-277) returns 1.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace psc {

// flag bits of SampleReads::flags (one byte per read)
enum : uint8_t {
	RF_PRIMARY = 1,        // (flag & 0x900) == 0            alignments.hpp:394-400
	RF_CLIP_HEAD = 2,      // first CIGAR op is S/H/P        alignments.hpp:210-217
	RF_CLIP_TAIL = 4,      // last CIGAR op is S/H/P
	RF_FIRST_MATE = 8,     // flag & 0x40                    alignments.hpp:356-361
	RF_UNIQUE = 16,        // !(NH > 1)                      alignments.hpp:382-392
	RF_MATE_SAME_CHR = 32  // mate mapped and mtid == tid    alignments.hpp:311-323, Constraints.cpp:337
};

struct SampleReads {
	std::string path;
	std::vector<std::string> chr_names;
	std::vector<int32_t> chr_len;
	// reads that pass Alignments::Next(), in file order (coordinate sorted)
	std::vector<int64_t> chr_off;   // [n_chr + 1] first read of every chromosome
	std::vector<int32_t> pos;       // segments[0].a (0-based)
	std::vector<int32_t> mate_pos;  // core.mpos
	std::vector<uint8_t> flags;     // RF_*
	std::vector<uint8_t> nh;        // NH tag, 1 if absent          alignments.hpp:448-455
	std::vector<uint64_t> name_key; // hash of the read id the mate lookup uses (see bam_io.cpp)
	std::vector<uint32_t> seg_off;  // [n_reads + 1]
	std::vector<int32_t> seg_a, seg_b; // aligned blocks, 0-based inclusive   alignments.hpp:200-250
	std::vector<uint32_t> prim_cum; // #primary records read up to and including this read (Q8)
	uint32_t total_primary = 0;     // #primary records in the whole file
	// Alignments::GetGeneralInfo(true)
	int32_t read_len = 0, frag_len = 0, frag_stdev = 0;
	int32_t mate_paired = 0;

	int64_t n_reads() const { return (int64_t)pos.size(); }
};

// The per-record filter shared by the BAM decoder and the synthetic generator: feed it every
// record of a coordinate-sorted stream in file order (out.chr_names must be set first).
struct ReadAccumulator {
	struct State;
	SampleReads &out;
	State *st;
	explicit ReadAccumulator(SampleReads &o);
	~ReadAccumulator();
	void add(int32_t tid, int32_t pos, uint32_t flag, int32_t l_seq, int32_t mtid, int32_t mpos,
	         const char *qname, size_t qname_len, const uint32_t *cigar, uint32_t n_cigar,
	         bool has_nh, int32_t nh);
	void finish();
};

// Decode a whole BAM.  `threads` BGZF inflate workers (>=1).  Returns false (and sets err) on
// malformed input.
bool decode_bam(const std::string &path, int threads, SampleReads &out, std::string &err);

// Read only the header (chromosome names, classes.cpp:269 uses BAM #0's).
bool read_bam_header(const std::string &path, std::vector<std::string> &names,
                     std::vector<int32_t> &lens, std::string &err);

// 64-bit hash of a read id (FNV-1a folded through a finaliser).
uint64_t hash_read_id(const char *s, size_t n);

// ---- writer (used by the synthetic generator and tests only) ----
struct BamWriter {
	void *impl = nullptr;
	bool open(const std::string &path, const std::vector<std::string> &names,
	          const std::vector<int32_t> &lens);
	// cigar: array of (len<<4 | op).  aux: raw aux bytes (may be empty).  seq is written as
	// l_seq 'A' bases (content is irrelevant to stage 3: alignments.hpp:475-487 / Q1).
	void write(int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, const std::string &name,
	           const uint32_t *cigar, int n_cigar, int32_t l_seq, int32_t mtid, int32_t mpos,
	           int32_t tlen, const std::string &aux);
	void close();
};

} // namespace psc
