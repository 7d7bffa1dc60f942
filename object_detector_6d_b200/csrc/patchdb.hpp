// The reference's on-disk patch database: an LMDB environment whose main database maps
// "%04d_%08d" (object, patch) keys to serialised caffe::Datum messages, written by patch_generator
// (PatchGen/src/patch_generator.cpp:382-523, 566-580) and read back by train_patch_generator
// (PatchGen/src/train_patch_generator.cpp:33-52, 79-101) and by Caffe's DATA layer when the auto-encoder is trained.
//
// liblmdb and libprotobuf are not in this image (and would be one more dependency of a drop-in), so this file reads and
// writes the two formats directly -- host code only, nothing here touches the GPU:
//
//   * LMDB data file `data.mdb`, format version 1 (lmdb 0.9.x): 4096-byte pages; pages 0 and 1 are meta pages (the one with
//     the larger transaction id is current); the main database is a B+tree of branch / leaf / overflow pages.  The READER
//     walks whatever tree the current meta page points at, so it reads files liblmdb wrote (any number of transactions,
//     free-list pages are simply never reached).  The WRITER bulk-loads keys that arrive in ascending order -- the order
//     patch_generator's keys have by construction -- into a compact tree (what `mdb_copy -c` produces): full leaves left to
//     right, branch levels on top, empty free-list database, one committed transaction.
//   * caffe::Datum (caffe.proto: channels = 1, height = 2, width = 3, data = 4 (bytes), label = 5, float_data = 6,
//     encoded = 7): the five fields the reference sets, in field order, exactly as protobuf's C++ serialiser emits them.
//
// Page and node layout restated from lmdb's published format (mdb.c: MDB_page, MDB_node, MDB_meta, MDB_db); parity of the
// page layout is UNPINNED (no liblmdb here to open the files with): oracle/patchdb.py is a second, independent restatement
// in Python that must read what this writes and write what this reads; the Datum bytes are pinned to google.protobuf.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace patchdb {

constexpr uint32_t kPage = 4096, kHdr = 16, kNodeHdr = 8;
constexpr uint32_t kMagic = 0xBEEFC0DEu, kVersion = 1;
constexpr uint16_t P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08, P_LEAF2 = 0x20;
constexpr uint16_t F_BIGDATA = 0x01, F_SUBDATA = 0x02, F_DUPDATA = 0x04;
constexpr uint64_t kInvalid = ~0ull;
constexpr uint32_t kNodeMax = (((kPage - kHdr) / 2) & ~1u) - 2;  // me_nodemax: two keys must fit a page
constexpr uint32_t kMaxKey = 511;
constexpr uint64_t kMapSize = 1099511627776ull;  // mdb_env_set_mapsize(.., 1 TB), patch_generator.cpp:566

inline void put16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void put32(uint8_t* p, uint32_t v) { put16(p, v & 0xffff); put16(p + 2, v >> 16); }
inline void put64(uint8_t* p, uint64_t v) { put32(p, (uint32_t)v); put32(p + 4, (uint32_t)(v >> 32)); }
inline uint32_t get16(const uint8_t* p) { return p[0] | (uint32_t)p[1] << 8; }
inline uint32_t get32(const uint8_t* p) { return get16(p) | get16(p + 2) << 16; }
inline uint64_t get64(const uint8_t* p) { return get32(p) | (uint64_t)get32(p + 4) << 32; }

// ------------------------------------------------------------------------------------------------ caffe::Datum
struct Datum {
    int32_t channels = 0, height = 0, width = 0, label = 0;
    bool has_label = false, encoded = false;
    std::string data;
    std::vector<float> float_data;
};

inline void put_varint(std::string& out, uint64_t v) {
    while (v >= 0x80) { out.push_back((char)(v | 0x80)); v >>= 7; }
    out.push_back((char)v);
}

// What Datum::SerializeToString gives for a message with channels, height, width, data and label set (proto2: set fields in
// field-number order; int32 as a sign-extended varint).
inline std::string encode_datum(int channels, int height, int width, const uint8_t* data, size_t n, int label) {
    std::string out;
    out.reserve(n + 24);
    out.push_back(0x08); put_varint(out, (uint64_t)(int64_t)channels);
    out.push_back(0x10); put_varint(out, (uint64_t)(int64_t)height);
    out.push_back(0x18); put_varint(out, (uint64_t)(int64_t)width);
    out.push_back(0x22); put_varint(out, n);
    out.append(reinterpret_cast<const char*>(data), n);
    out.push_back(0x28); put_varint(out, (uint64_t)(int64_t)label);
    return out;
}

inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64 && p < end; shift += 7) {
        const uint8_t b = *p++;
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

// Any valid Datum: unknown fields skipped, float_data packed or not, later values of a scalar replace earlier ones.
inline bool decode_datum(const uint8_t* p, size_t n, Datum& d, std::string& err) {
    const uint8_t* end = p + n;
    d = Datum();
    while (p < end) {
        uint64_t tag, v;
        if (!get_varint(p, end, tag)) { err = "truncated field tag"; return false; }
        const uint32_t field = (uint32_t)(tag >> 3), wire = (uint32_t)(tag & 7);
        if (wire == 0) {
            if (!get_varint(p, end, v)) { err = "truncated varint"; return false; }
            if (field == 1) d.channels = (int32_t)v;
            else if (field == 2) d.height = (int32_t)v;
            else if (field == 3) d.width = (int32_t)v;
            else if (field == 5) { d.label = (int32_t)v; d.has_label = true; }
            else if (field == 7) d.encoded = v != 0;
        } else if (wire == 2) {
            if (!get_varint(p, end, v) || v > (uint64_t)(end - p)) { err = "truncated length-delimited field"; return false; }
            if (field == 4) d.data.assign(reinterpret_cast<const char*>(p), (size_t)v);
            else if (field == 6) {
                if (v % 4) { err = "packed float_data is not a multiple of 4 bytes"; return false; }
                for (uint64_t i = 0; i < v; i += 4) { float f; memcpy(&f, p + i, 4); d.float_data.push_back(f); }
            }
            p += v;
        } else if (wire == 5) {
            if (end - p < 4) { err = "truncated fixed32"; return false; }
            if (field == 6) { float f; memcpy(&f, p, 4); d.float_data.push_back(f); }
            p += 4;
        } else if (wire == 1) {
            if (end - p < 8) { err = "truncated fixed64"; return false; }
            p += 8;
        } else { err = "unsupported wire type " + std::to_string(wire); return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ writer
class Writer {
  public:
    ~Writer() { if (f_) fclose(f_); }
    bool open(const std::string& dir, std::string& err) {
        path_ = dir + "/data.mdb";
        if (FILE* probe = fopen(path_.c_str(), "rb")) {  // mdb_open on an existing environment would append; the reference's
            fclose(probe);                               // message asks "Does the lmdb already exist?" (patch_generator.cpp:573)
            err = "patch database " + path_ + " already exists";
            return false;
        }
        f_ = fopen(path_.c_str(), "wb");
        if (!f_) { err = "cannot create " + path_; return false; }
        uint8_t zero[2 * kPage] = {0};
        if (fwrite(zero, 1, sizeof zero, f_) != sizeof zero) { err = "cannot write " + path_; return false; }  // metas: at close
        next_pg_ = 2;
        begin_leaf();
        return true;
    }
    // Keys must arrive in strictly ascending byte order (MDB_APPEND semantics).
    bool put(const std::string& key, const void* data, size_t n, std::string& err) {
        if (!f_) { err = "patch database is not open"; return false; }
        if (key.empty() || key.size() > kMaxKey) { err = "key length " + std::to_string(key.size()) + " outside 1.." + std::to_string(kMaxKey); return false; }
        if (entries_ && key.compare(last_key_) <= 0) { err = "key '" + key + "' does not sort after '" + last_key_ + "'"; return false; }
        const bool big = kNodeHdr + key.size() + n > kNodeMax;
        const uint32_t node = (uint32_t)((kNodeHdr + key.size() + (big ? 8 : n) + 1) & ~(size_t)1);
        if (node + 2 > upper_ - lower_) {
            if (!flush_leaf(err)) return false;
            begin_leaf();
        }
        if (lower_ == kHdr) leaf_first_key_ = key;
        uint64_t ov_pg = 0;
        if (big) {  // the value goes to its own run of overflow pages, written right away; the leaf follows later
            const uint64_t pages = (kHdr - 1 + n) / kPage + 1;
            ov_pg = next_pg_;
            next_pg_ += pages;
            std::vector<uint8_t> ov((size_t)pages * kPage, 0);
            put64(&ov[0], ov_pg);
            put16(&ov[10], P_OVERFLOW);
            put32(&ov[12], (uint32_t)pages);
            memcpy(&ov[kHdr], data, n);
            if (!write_at(ov_pg, ov.data(), ov.size(), err)) return false;
            overflow_pages_ += pages;
        }
        upper_ -= node;
        uint8_t* nd = &page_[upper_];
        put16(nd, (uint32_t)(n & 0xffff));
        put16(nd + 2, (uint32_t)(n >> 16));
        put16(nd + 4, big ? F_BIGDATA : 0);
        put16(nd + 6, (uint32_t)key.size());
        memcpy(nd + kNodeHdr, key.data(), key.size());
        if (big) put64(nd + kNodeHdr + key.size(), ov_pg);
        else memcpy(nd + kNodeHdr + key.size(), data, n);
        put16(&page_[lower_], upper_);
        lower_ += 2;
        ++entries_;
        last_key_ = key;
        return true;
    }
    bool close(std::string& err) {
        if (!f_) { err = "patch database is not open"; return false; }
        uint64_t root = kInvalid, branch_pages = 0;
        uint32_t depth = 0;
        if (entries_) {
            if (!flush_leaf(err)) return false;
            depth = 1;
            std::vector<std::pair<std::string, uint64_t>> level;
            level.swap(children_);
            while (level.size() > 1) {  // one branch level above `level`
                std::vector<std::pair<std::string, uint64_t>> up;
                size_t i = 0;
                while (i < level.size()) {
                    std::vector<uint8_t> pg(kPage, 0);
                    uint32_t lower = kHdr, upper = kPage;
                    const size_t first = i;
                    for (; i < level.size(); ++i) {
                        const std::string key = i == first ? std::string() : level[i].first;  // leftmost key of a branch page is empty
                        const uint32_t node = (uint32_t)((kNodeHdr + key.size() + 1) & ~(size_t)1);
                        if (node + 2 > upper - lower) break;
                        upper -= node;
                        uint8_t* nd = &pg[upper];
                        const uint64_t child = level[i].second;
                        put16(nd, (uint32_t)(child & 0xffff));
                        put16(nd + 2, (uint32_t)(child >> 16 & 0xffff));
                        put16(nd + 4, (uint32_t)(child >> 32 & 0xffff));
                        put16(nd + 6, (uint32_t)key.size());
                        memcpy(nd + kNodeHdr, key.data(), key.size());
                        put16(&pg[lower], upper);
                        lower += 2;
                    }
                    // a branch page needs two children: never leave a single child for the last page of a level
                    if (i + 1 == level.size() && i - first > 2) {
                        --i;  // hand the last fitted child to the next page
                        const uint32_t back = get16(&pg[lower - 2]);
                        const uint32_t ksz = get16(&pg[back + 6]);
                        memset(&pg[back], 0, (kNodeHdr + ksz + 1) & ~1u);
                        upper = back + ((kNodeHdr + ksz + 1) & ~1u);
                        lower -= 2;
                        put16(&pg[lower], 0);
                    }
                    const uint64_t pgno = next_pg_++;
                    put64(&pg[0], pgno);
                    put16(&pg[10], P_BRANCH);
                    put16(&pg[12], lower);
                    put16(&pg[14], upper);
                    if (!write_at(pgno, pg.data(), kPage, err)) return false;
                    ++branch_pages;
                    up.emplace_back(level[first].first, pgno);
                }
                level.swap(up);
                ++depth;
            }
            root = level[0].second;
        }
        // both meta pages: page 0 holds transaction 0 (the empty environment mdb_env_open creates), page 1 transaction 1
        for (int m = 0; m < 2; ++m) {
            uint8_t pg[kPage] = {0};
            put64(pg, (uint64_t)m);
            put16(pg + 10, P_META);
            uint8_t* mm = pg + kHdr;
            put32(mm, kMagic);
            put32(mm + 4, kVersion);
            put64(mm + 8, 0);             // mm_address
            put64(mm + 16, kMapSize);
            uint8_t* free_db = mm + 24;   // MDB_db: pad u32, flags u16, depth u16, branch, leaf, overflow, entries, root
            put32(free_db, kPage);        // mm_psize
            put16(free_db + 4, 0x08);     // MDB_INTEGERKEY
            put64(free_db + 40, kInvalid);
            uint8_t* main_db = mm + 72;
            if (m == 1 && entries_) {
                put16(main_db + 6, depth);
                put64(main_db + 8, branch_pages);
                put64(main_db + 16, leaf_pages_);
                put64(main_db + 24, overflow_pages_);
                put64(main_db + 32, entries_);
                put64(main_db + 40, root);
            } else put64(main_db + 40, kInvalid);
            put64(mm + 120, m == 1 && entries_ ? next_pg_ - 1 : 1);  // mm_last_pg
            put64(mm + 128, (uint64_t)(m == 1 && entries_ ? 1 : 0));  // mm_txnid
            if (!write_at((uint64_t)m, pg, kPage, err)) return false;
        }
        const bool ok = fflush(f_) == 0;
        fclose(f_);
        f_ = nullptr;
        if (!ok) err = "cannot write " + path_;
        return ok;
    }
    uint64_t entries() const { return entries_; }

  private:
    void begin_leaf() {
        memset(page_, 0, sizeof page_);
        lower_ = kHdr;
        upper_ = kPage;
    }
    bool flush_leaf(std::string& err) {
        const uint64_t pgno = next_pg_++;
        put64(page_, pgno);
        put16(page_ + 10, P_LEAF);
        put16(page_ + 12, lower_);
        put16(page_ + 14, upper_);
        if (!write_at(pgno, page_, kPage, err)) return false;
        children_.emplace_back(leaf_first_key_, pgno);
        ++leaf_pages_;
        return true;
    }
    bool write_at(uint64_t pgno, const uint8_t* p, size_t n, std::string& err) {
        if (fseeko(f_, (off_t)(pgno * kPage), SEEK_SET) != 0 || fwrite(p, 1, n, f_) != n) { err = "cannot write " + path_; return false; }
        return true;
    }
    FILE* f_ = nullptr;
    std::string path_, last_key_, leaf_first_key_;
    uint8_t page_[kPage];
    uint32_t lower_ = kHdr, upper_ = kPage;
    uint64_t next_pg_ = 2, entries_ = 0, leaf_pages_ = 0, overflow_pages_ = 0;
    std::vector<std::pair<std::string, uint64_t>> children_;
};

// ------------------------------------------------------------------------------------------------ reader
class Reader {
  public:
    bool open(const std::string& dir, std::string& err) {
        const std::string path = dir + "/data.mdb";
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) { err = "cannot open " + path; return false; }
        fseeko(f, 0, SEEK_END);
        const off_t size = ftello(f);
        fseeko(f, 0, SEEK_SET);
        file_.resize((size_t)size);
        const bool ok = size >= (off_t)(2 * kPage) && fread(file_.data(), 1, file_.size(), f) == file_.size();
        fclose(f);
        if (!ok) { err = path + " is shorter than its two meta pages"; return false; }
        int best = -1;
        uint64_t best_txn = 0;
        for (int m = 0; m < 2; ++m) {
            const uint8_t* pg = &file_[(size_t)m * kPage];
            const uint8_t* mm = pg + kHdr;
            if (!(get16(pg + 10) & P_META) || get32(mm) != kMagic) continue;
            if (get32(mm + 4) != kVersion) { err = path + ": data format version " + std::to_string(get32(mm + 4)) + ", this reader knows 1"; return false; }
            const uint64_t txn = get64(mm + 128);
            if (best < 0 || txn > best_txn) { best = m; best_txn = txn; }
        }
        if (best < 0) { err = path + " has no valid LMDB meta page"; return false; }
        const uint8_t* mm = &file_[(size_t)best * kPage + kHdr];
        if (get32(mm + 24) != kPage) { err = path + ": page size " + std::to_string(get32(mm + 24)) + ", this reader knows 4096"; return false; }
        const uint8_t* main_db = mm + 72;
        if (get16(main_db + 4) != 0) { err = path + ": main database has flags (DUPSORT / INTEGERKEY / ...), patch databases do not"; return false; }
        entries_ = get64(main_db + 32);
        root_ = get64(main_db + 40);
        last_pg_ = get64(mm + 120);
        if (root_ != kInvalid && ((root_ + 1) * kPage > file_.size() || root_ > last_pg_)) { err = path + ": root page lies outside the file"; return false; }
        stack_.clear();
        started_ = false;
        return true;
    }
    uint64_t entries() const { return entries_; }
    // Entries in key order (MDB_FIRST, MDB_NEXT ..).  false with an empty `err` = end of the database.
    bool next(std::string& key, std::string& value, std::string& err) {
        err.clear();
        if (!started_) {
            started_ = true;
            if (root_ == kInvalid) return false;
            if (!descend(root_, err)) return false;
        } else {
            while (!stack_.empty()) {  // advance the cursor: next node of the leaf, else climb and take the next branch
                Level& top = stack_.back();
                if (++top.idx < top.n) {
                    if (top.leaf) break;
                    const uint64_t child = branch_child(top);
                    if (!descend(child, err)) return false;
                    break;
                }
                stack_.pop_back();
            }
            if (stack_.empty()) return false;
        }
        const Level& lf = stack_.back();
        const uint8_t* pg = &file_[(size_t)lf.pgno * kPage];
        const uint32_t off = get16(pg + kHdr + 2 * lf.idx);
        if (off < kHdr + 2 * lf.n || off + kNodeHdr > kPage) { err = "leaf page " + std::to_string(lf.pgno) + ": node offset outside the page"; return false; }
        const uint8_t* nd = pg + off;
        const uint32_t dsize = get16(nd) | get16(nd + 2) << 16, flags = get16(nd + 4), ksize = get16(nd + 6);
        if (flags & (F_SUBDATA | F_DUPDATA)) { err = "sub-databases / duplicate keys are not patch-database features"; return false; }
        if (off + kNodeHdr + ksize > kPage) { err = "leaf page " + std::to_string(lf.pgno) + ": key outside the page"; return false; }
        key.assign(reinterpret_cast<const char*>(nd + kNodeHdr), ksize);
        if (flags & F_BIGDATA) {
            if (off + kNodeHdr + ksize + 8 > kPage) { err = "leaf page: overflow reference outside the page"; return false; }
            const uint64_t ov = get64(nd + kNodeHdr + ksize);
            if (ov > last_pg_ || ov * kPage + kHdr + dsize > file_.size()) { err = "overflow page " + std::to_string(ov) + " lies outside the file"; return false; }
            const uint8_t* op = &file_[(size_t)ov * kPage];
            if (!(get16(op + 10) & P_OVERFLOW)) { err = "page " + std::to_string(ov) + " is not an overflow page"; return false; }
            value.assign(reinterpret_cast<const char*>(op + kHdr), dsize);
        } else {
            if (off + kNodeHdr + ksize + dsize > kPage) { err = "leaf page " + std::to_string(lf.pgno) + ": value outside the page"; return false; }
            value.assign(reinterpret_cast<const char*>(nd + kNodeHdr + ksize), dsize);
        }
        return true;
    }

  private:
    struct Level { uint64_t pgno; uint32_t idx, n; bool leaf; };
    uint64_t branch_child(const Level& l) const {
        const uint8_t* pg = &file_[(size_t)l.pgno * kPage];
        const uint8_t* nd = pg + get16(pg + kHdr + 2 * l.idx);
        return get16(nd) | (uint64_t)get16(nd + 2) << 16 | (uint64_t)get16(nd + 4) << 32;
    }
    bool descend(uint64_t pgno, std::string& err) {  // leftmost leaf below pgno
        for (;;) {
            if (pgno > last_pg_ || (pgno + 1) * kPage > file_.size()) { err = "page " + std::to_string(pgno) + " lies outside the file"; return false; }
            if (stack_.size() > 64) { err = "B-tree deeper than 64 levels (a cycle)"; return false; }
            const uint8_t* pg = &file_[(size_t)pgno * kPage];
            const uint32_t flags = get16(pg + 10), lower = get16(pg + 12);
            if (get64(pg) != pgno) { err = "page " + std::to_string(pgno) + " carries page number " + std::to_string(get64(pg)); return false; }
            if (flags & P_LEAF2) { err = "fixed-size-key leaf pages are not a patch-database feature"; return false; }
            if (lower < kHdr || lower > kPage) { err = "page " + std::to_string(pgno) + ": bad lower bound"; return false; }
            const uint32_t n = (lower - kHdr) / 2;
            if (!n) { err = "page " + std::to_string(pgno) + " is empty"; return false; }
            if (flags & P_LEAF) { stack_.push_back(Level{pgno, 0, n, true}); return true; }
            if (!(flags & P_BRANCH)) { err = "page " + std::to_string(pgno) + " is neither branch nor leaf"; return false; }
            stack_.push_back(Level{pgno, 0, n, false});
            const uint32_t off = get16(pg + kHdr);
            if (off + kNodeHdr > kPage) { err = "branch page " + std::to_string(pgno) + ": node offset outside the page"; return false; }
            pgno = branch_child(stack_.back());
        }
    }
    std::vector<uint8_t> file_;
    std::vector<Level> stack_;
    uint64_t entries_ = 0, root_ = kInvalid, last_pg_ = 0;
    bool started_ = false;
};

}  // namespace patchdb
