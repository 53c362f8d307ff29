// zpq_host.h -- internal host-side declarations of libzpaqb200 (not part of the C ABI).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zpaqb200.h"
#include "zpq_plan.h"

namespace zpq {

typedef std::vector<uint8_t> Bytes;

struct Failure : std::runtime_error {
  int code;
  Failure(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- front end (zpq_frontend.cpp) ----
int comp_len(int type);
int block_arg0(uint64_t n);
std::string expand_method(const std::string& method, const uint8_t* data, uint64_t n);
constexpr int kGapBins = 1 << 12;        // NR, LibZPAQ.cs:242
bool method_needs_analysis(const std::string& method);      // levels 5..9 look at the data (LibZPAQ.cs:242-280)
void gap_histogram(const uint8_t* data, uint64_t n, int* gap /*kGapBins*/);
std::string expand_method_gaps(const std::string& method, uint64_t n, const int* gaps /*kGapBins, or null for levels 0..4*/);
std::string make_config(const std::string& method, int args[9]);
void compile_config(const std::string& text, const int* args, Bytes& hdr, Bytes& pcomp, std::string* pcomp_cmd);
void builtin_model(int level, Bytes& hdr);
struct PostCandidate { uint32_t kind, e8, param; int wild; Bytes prog; };
void post_candidates(int ph, int pm, std::vector<PostCandidate>& out);

// ---- model planning (zpq_model.cpp) ----
struct Header {           // a parsed block header (ZPAQL.read, ZPAQL.cs:112-156)
  Bytes wire;             // hsize.. COMP 0 HCOMP 0, as stored in the archive
  int n;
  int cend;               // end of COMP incl. its END byte, in wire[]
  int hh, hm, ph, pm;
};
// Parse a header from p[0..avail); returns bytes consumed.  Throws Failure(ZPQ_E_CORRUPT).
size_t parse_header(const uint8_t* p, size_t avail, Header& h);
double header_memory(const Header& h);  // ZPAQL.memory(), ZPAQL.cs:58-81
// Build the device plan.  smem_budget = shared bytes one resident block may use for its slice.
// duo_g != 0 lays the shared slice out for the two-role encoder (zpq_duo.cuh) with duo_g lanes per block.
// fdec lays it out for the speculative decoder (zpq_fdec.cuh).
// max_stream != 0: no stream coded with this plan is longer (lets a MATCH buffer be allocated at the size the streams can reach).
void build_plan(const Header& h, bool for_decode, uint32_t smem_budget, Plan& plan, int duo_g = 0, bool fdec = false, uint64_t max_stream = 0);
void build_tables(Tables& t);           // throws if the squash/stretch checksums are off
void sha1_host(const uint8_t* p, uint64_t n, uint8_t out[20]);

// ---- archive framing on the host (zpq_model.cpp) ----
extern const uint8_t kLocatorTag[13];   // Compressor.cs:27-43
struct SegmentRef {
  uint64_t data_off, data_len;   // coded bytes incl. the 4 zero bytes that end them
  bool has_sha1;
  uint8_t sha1[20];
  int64_t size_hint;             // decimal size from the comment, or -1
};
struct BlockRef {
  Header hdr;
  std::vector<SegmentRef> segs;
  uint64_t end;                  // one past the 255 byte, relative to the block start given
};
// Parse one archive block that starts at its tag or at "zPQ" (Decompresser.cs:29-108,163-194;
// segment ends located like Decoder.skip, Decoder.cs:70-98).
void parse_block(const uint8_t* p, uint64_t avail, BlockRef& out);
int64_t find_blocks(const uint8_t* p, uint64_t n, uint64_t* offsets, uint64_t max_blocks);

}  // namespace zpq
