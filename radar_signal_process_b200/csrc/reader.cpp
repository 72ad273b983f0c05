// reader.cpp -- wire-format frame parser and cross-file byte stream.  Host code.
//
// Replaces read_continuous_file_stream.m:22-168 (one logical byte stream over the numbered capture files
// 1.00000N.bin, DataFullPathGen.m:10-16) and the per-PRT parser of FrameDataRead_xzr.m:57-198 for DDC data:
//   64 B head (16 x uint32) | 128 B realtime block | payload (n * ch * 4 B, padded to 64 B) | 64 B tail.
// It feeds the int16 payloads, PRT after PRT, straight into the caller's (pinned) batch buffer in the layout
// rb200_chain_i16 consumes, so a capture can go disk -> pinned host memory -> GPU without ever becoming a
// MATLAB matrix.  The reference's file-index behaviour is reproduced exactly, including the double increment
// after a read that ends exactly at the end of a file (read_continuous_file_stream.m:148 then :48), which
// skips the next file; RB200_READER_NO_SKIP_QUIRK=1 switches that off.
#include <algorithm>
#include <cstdint>
#include <exception>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/radar_b200.h"

struct rb200_reader {
    std::string dir;
    bool is_open = false;
    FILE* f = nullptr;
    long long pos = 0, max_len = 0;
    int file_index = 0;                 // starts at 0, incremented before the first open (…m:43,48)
    bool skip_quirk = true;
    std::string err;
    std::vector<uint8_t> scratch;
};

static bool dir_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

// DataFullPathGen.m:10-27
static std::string file_path(const rb200_reader* r, int idx) {
    char name[64];
    if (idx < 10) snprintf(name, sizeof name, "1.00000%d.bin", idx);
    else if (idx < 100) snprintf(name, sizeof name, "1.0000%d.bin", idx);
    else snprintf(name, sizeof name, "1.000%d.bin", idx);
    const std::string sub = r->dir + "/\xE9\x9B\xB7\xE8\xBE\xBE\xE5\x8E\x9F\xE5\xA7\x8B\xE6\x95\xB0\xE6\x8D\xAE";   // "雷达原始数据"
    return (dir_exists(sub) ? sub : r->dir) + "/" + name;
}

static bool open_index(rb200_reader* r, int idx) {
    r->f = fopen(file_path(r, idx).c_str(), "rb");
    if (!r->f) return false;
    fseek(r->f, 0, SEEK_END);
    r->max_len = ftell(r->f);
    fseek(r->f, 0, SEEK_SET);
    r->pos = 0;
    r->is_open = true;
    return true;
}

// read_continuous_file_stream: returns bytes actually read; *eos set like is_end_of_stream
static long long stream_read(rb200_reader* r, uint8_t* dst, long long want, bool* eos) {
    *eos = false;
    long long got = 0;
    if (!r->is_open) {
        r->file_index += 1;                                           // :48
        if (!open_index(r, r->file_index)) {
            r->is_open = false;
            *eos = true;                                              // :55-59
            return 0;
        }
    }
    if (r->pos + want > r->max_len) {                                 // :85 read spans the end of the file
        long long part = r->max_len - r->pos;
        if (part < 0) part = 0;
        got = (long long)fread(dst, 1, (size_t)part, r->f);
        fclose(r->f);
        r->f = nullptr;
        r->is_open = false;
        const long long remain = want - got;
        if (remain > 0) {
            r->file_index += 1;                                       // :101
            if (!open_index(r, r->file_index)) {
                *eos = true;                                          // :106-113
                r->pos = 0;
                r->max_len = 0;
                return got;
            }
            const long long g2 = (long long)fread(dst + got, 1, (size_t)remain, r->f);
            got += g2;
            r->pos += g2;                                             // :133
        }
    } else if (r->pos + want == r->max_len) {                         // :137 read ends exactly at the end of the file
        got = (long long)fread(dst, 1, (size_t)want, r->f);
        fclose(r->f);
        r->f = nullptr;
        r->is_open = false;
        if (r->skip_quirk) r->file_index += 1;                        // :148 (and :48 increments again on the next call)
        r->pos = 0;
        r->max_len = 0;
    } else {
        got = (long long)fread(dst, 1, (size_t)want, r->f);
        r->pos += got;
    }
    if (got < want && r->is_open) *eos = true;                        // :160-163
    return got;
}

extern "C" int rb200_reader_open(rb200_reader** out, const char* dir) {
    if (!out || !dir) return RB200_ERR_ARG;
    *out = nullptr;
    if (!dir_exists(dir)) return RB200_ERR_ARG;                       // DataFullPathGen.m:5-7 raises
    rb200_reader* r = new rb200_reader();
    r->dir = dir;
    const char* q = getenv("RB200_READER_NO_SKIP_QUIRK");
    r->skip_quirk = !(q && atoi(q));
    *out = r;
    return RB200_OK;
}

extern "C" int rb200_reader_close(rb200_reader* r) {
    if (!r) return RB200_OK;
    if (r->f) fclose(r->f);
    delete r;
    return RB200_OK;
}

extern "C" const char* rb200_reader_last_error(const rb200_reader* r) { return r ? r->err.c_str() : ""; }

extern "C" int rb200_reader_state(const rb200_reader* r, int* file_index, long long* pos) {
    if (!r) return RB200_ERR_ARG;
    if (file_index) *file_index = r->file_index;
    if (pos) *pos = r->pos;
    return RB200_OK;
}

// One logical frame of n_prt PRTs of the wanted data type (FrameDataRead_xzr.m:57-198).  DDC (data_type 1): out receives
// the payload only, [prt][range][channel][I,Q] int16.  DBF-type (data_type 2): out receives every PRT's payload including
// its padding to 64 bytes, exactly the layout rb200_chain_dbf24 / rb200_unpack_dbf24 take.  Per-PRT header fields go to the
// optional arrays.
static int next_frame(rb200_reader* r, int want_type, int n_prt, int n_range, int n_channels, uint8_t* out, uint32_t* frame_no,
                      uint16_t* servo_angle, uint64_t* timer_cnt, int* prts_read, int* end_of_stream) {
    if (!r || !out || n_prt < 1 || n_range < 1 || n_channels < 1 || !prts_read || !end_of_stream) return RB200_ERR_ARG;
    *prts_read = 0;
    *end_of_stream = 0;
    const long long head_b = 64, rt_b = 128, tail_b = 64;             // bin_to_mat_xzr.m:41-43
    uint8_t head[64];
    std::vector<uint8_t>& buf = r->scratch;
    for (int prt = 0; prt < n_prt; ++prt) {
        bool eos;
        if (stream_read(r, head, head_b, &eos) < head_b || eos) { *end_of_stream = 1; return RB200_OK; }        // :62-67
        uint32_t h[16];
        memcpy(h, head, 64);                                          // :70 typecast uint32 (little-endian host)
        const uint32_t channel_num = h[3] & 0xFFu;                    // :77
        const uint32_t pulse_data_num = h[6];                         // :79
        const uint32_t data_type = h[7] & 0xFFu;                      // :80
        if (pulse_data_num == 0) { *end_of_stream = 1; r->err = "invalid pulse_data_num in the PRT head"; return RB200_OK; }   // :90-94
        uint8_t rt[128];
        if (stream_read(r, rt, rt_b, &eos) < rt_b || eos) { *end_of_stream = 1; return RB200_OK; }               // :97-102
        long long sig;                                                // :105-113
        if (data_type == 0) sig = (long long)pulse_data_num * channel_num * 2;
        else if (data_type == 1) sig = (long long)pulse_data_num * channel_num * 2 * 2;
        else sig = (long long)pulse_data_num * channel_num * 6 + (long long)pulse_data_num * (8 - (6 * channel_num) % 8);
        const long long padded = sig + ((sig % 64) ? 64 - sig % 64 : 0);                                         // :115-119
        // The head is unvalidated capture data: a corrupt pulse_data_num / channel_num must not turn into a terabyte
        // allocation.  A payload far beyond the configured geometry can only end in the short read the reference gets
        // from fread (:122-127), so report end-of-stream without allocating.
        const long long expect = (long long)n_range * n_channels * 8 + 64;
        if (padded > std::max<long long>(4 * expect, 1ll << 20)) {
            *end_of_stream = 1;
            r->err = "PRT head announces a payload far larger than the configured geometry (corrupt capture?)";
            return RB200_OK;
        }
        buf.resize((size_t)padded);
        if (stream_read(r, buf.data(), padded, &eos) < padded || eos) { *end_of_stream = 1; return RB200_OK; }   // :122-127
        if ((int)data_type != want_type) {
            r->err = want_type == 1 ? "rb200_reader_next_frame_ddc: PRT is not DDC data (data_type != 1)"
                                    : "rb200_reader_next_frame_dbf24: PRT is not DBF-type data (data_type != 2)";
            return RB200_ERR_UNSUPPORTED;
        }
        // :171-176 size check of the parsed PRT against the configured geometry
        if ((int)pulse_data_num != n_range || (int)channel_num != n_channels) {
            *end_of_stream = 1;
            r->err = "PRT geometry differs from the configured point_PRT / channel_num";
            return RB200_OK;
        }
        if (want_type == 1) memcpy(out + (size_t)prt * n_range * n_channels * 4, buf.data(), (size_t)n_range * n_channels * 4);   // :138,150
        else memcpy(out + (size_t)prt * (size_t)padded, buf.data(), (size_t)padded);                               // :130-135 decode on the device
        if (frame_no) frame_no[prt] = h[0];                           // :74
        if (servo_angle) servo_angle[prt] = (uint16_t)(h[4] & 0xFFFFu);   // :78
        if (timer_cnt) timer_cnt[prt] = (uint64_t)h[8] + ((uint64_t)h[9] << 32);   // :83
        *prts_read = prt + 1;                                         // :179
        uint8_t tail[64];
        if (stream_read(r, tail, tail_b, &eos) < tail_b || eos) { *end_of_stream = 1; return RB200_OK; }          // :184-189
    }
    return RB200_OK;
}

extern "C" int rb200_reader_next_frame_ddc(rb200_reader* r, int n_prt, int n_range, int n_channels, int16_t* raw_out,
                                           uint32_t* frame_no, uint16_t* servo_angle, uint64_t* timer_cnt,
                                           int* prts_read, int* end_of_stream) {
    try {
        return next_frame(r, 1, n_prt, n_range, n_channels, reinterpret_cast<uint8_t*>(raw_out), frame_no, servo_angle, timer_cnt, prts_read,
                          end_of_stream);
    } catch (const std::exception& e) {          // no C++ exception crosses the C ABI
        if (r) r->err = std::string("rb200_reader_next_frame_ddc: ") + e.what();
        if (end_of_stream) *end_of_stream = 1;
        return RB200_ERR_ARG;
    }
}

extern "C" int rb200_reader_next_frame_dbf24(rb200_reader* r, int n_prt, int n_range, int n_channels, uint8_t* payload_out,
                                             uint32_t* frame_no, uint16_t* servo_angle, uint64_t* timer_cnt,
                                             int* prts_read, int* end_of_stream) {
    try {
        return next_frame(r, 2, n_prt, n_range, n_channels, payload_out, frame_no, servo_angle, timer_cnt, prts_read, end_of_stream);
    } catch (const std::exception& e) {
        if (r) r->err = std::string("rb200_reader_next_frame_dbf24: ") + e.what();
        if (end_of_stream) *end_of_stream = 1;
        return RB200_ERR_ARG;
    }
}
