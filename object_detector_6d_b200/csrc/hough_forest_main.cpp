// HoughForest -- command-line drop-in for the reference's `HoughForest --test` (HoughForest/src/main.cpp:33-78 and
// HFTest::DetectObjects, HoughForest/src/HFTest.cpp:1152-1311), on top of the C ABI in include/hf6d.h.
//
//   HoughForest --test --detector_options_file=<options.txt> [--output_folder=<dir>]  < pairs.txt
//
// stdin carries whitespace-separated `rgb_path depth_path` pairs until EOF (HFTest.cpp:1238).  For every pair the
// program writes `<output_folder><stem>_res.txt` (object name, instance counter, 4x4 pose in Eigen's default stream
// format, blank line -- HFTest.cpp:1264-1290) and `<output_folder><stem>_res.png`, and prints the reference's progress
// lines to stdout.  Unreadable images are reported and skipped (HFTest.cpp:1242-1251).
//
// With the objects' mesh files in place every hypothesis goes through ICP, hypothesis scoring and the joint
// optimisation on the GPU (hf6d_refine) and the poses written are the refined ones of the selected hypotheses, in final-score
// order, at most `instances` per object, with MeshUtils::renderObject's overlay on the result image -- the reference's output.
// Without the meshes (the reference aborts there) the *pre-ICP* poses (HFTest.cpp:922-924 + MeshUtils.cpp:423-440) are written,
// ranked by the Hough terms of the reference's final score, pose_score * pose_score_coeff + location_score *
// location_score_coeff (MeshUtils.cpp:780-784), and the program says so.
//
//   HoughForest --train --input=<training vectors> --output=<dir> --patch_size_in_voxels=8 --voxel_size_in_m=0.005 [...]
//
// trains a forest on the GPU (main.cpp:41-66, HFTrain::train) and writes forest.txt + tree<N>.dat.
//
// Host code only: every frame goes through hf6d_submit / hf6d_wait (+ hf6d_refine), training through hf6d_train_forest; there
// is no CPU detection or training path in this file.
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hf6d.h"

namespace {

// ------------------------------------------------------------------------------------------------ image files
struct Image {
    int w = 0, h = 0, channels = 0, bits = 0;  // bits per sample: 8 or 16
    std::vector<uint16_t> px;                  // samples widened to 16 bits, interleaved
};

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    if (n < 0) return false;
    f.seekg(0);
    out.resize((size_t)n);
    if (n) f.read(reinterpret_cast<char*>(out.data()), n);
    return (bool)f;
}

uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Non-interlaced PNG, colour types 0/2/3/4/6, bit depths 1..16 (what cv::imread accepts for these inputs).
bool decode_png(const std::vector<uint8_t>& file, Image& img, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8)) { err = "not a PNG file"; return false; }
    size_t at = 8;
    int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    bool end = false;
    while (!end && at + 12 <= file.size()) {
        const uint32_t len = be32(&file[at]);
        const char* type = reinterpret_cast<const char*>(&file[at + 4]);
        if (at + 12 + (size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
        const uint8_t* body = &file[at + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        at += 12 + (size_t)len;
    }
    if (w <= 0 || h <= 0 || w > 65535 || h > 65535) { err = "bad PNG dimensions"; return false; }
    if (interlace) { err = "interlaced PNG is not supported"; return false; }
    int nch;
    switch (ctype) {
        case 0: nch = 1; break;
        case 2: nch = 3; break;
        case 3: nch = 1; break;
        case 4: nch = 2; break;
        case 6: nch = 4; break;
        default: err = "bad PNG colour type"; return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "bad PNG bit depth"; return false; }
    const size_t bpp_bits = (size_t)nch * depth, stride = ((size_t)w * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    if (idat.empty() || uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) {
        err = "PNG data does not inflate to the image size";
        return false;
    }
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    const bool pal = ctype == 3;
    const int out_ch = pal ? 3 : nch;
    img.w = w; img.h = h; img.channels = out_ch; img.bits = (depth == 16) ? 16 : 8;
    img.px.assign((size_t)w * h * out_ch, 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(stride + 1) * (size_t)y];
        const int filter = line[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = line[1 + i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: err = "bad PNG filter"; return false;
            }
            cur[i] = (uint8_t)v;
        }
        uint16_t* dst = &img.px[(size_t)y * w * out_ch];
        for (int x = 0; x < w; ++x)
            for (int ch = 0; ch < nch; ++ch) {
                const size_t s = (size_t)x * nch + ch;
                unsigned v;
                if (depth == 16) v = (unsigned)cur[2 * s] << 8 | cur[2 * s + 1];
                else if (depth == 8) v = cur[s];
                else {
                    const size_t bit = s * depth;
                    v = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1u << depth) - 1);
                    if (!pal) v = v * 255u / ((1u << depth) - 1);  // grey samples scale to 8 bits
                }
                if (pal) {
                    if ((size_t)v * 3 + 2 >= palette.size()) { err = "PNG palette index out of range"; return false; }
                    dst[x * 3 + 0] = palette[v * 3]; dst[x * 3 + 1] = palette[v * 3 + 1]; dst[x * 3 + 2] = palette[v * 3 + 2];
                } else dst[(size_t)x * nch + ch] = (uint16_t)v;
            }
        prev.swap(cur);
    }
    return true;
}

// Binary PGM / PPM (P5 / P6), 8 or 16 bit big-endian samples.
bool decode_pnm(const std::vector<uint8_t>& file, Image& img, std::string& err) {
    if (file.size() < 3 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) { err = "not a binary PGM/PPM file"; return false; }
    size_t at = 2;
    auto next_int = [&](long& v) {
        for (;;) {
            while (at < file.size() && isspace(file[at])) ++at;
            if (at < file.size() && file[at] == '#') { while (at < file.size() && file[at] != '\n') ++at; continue; }
            break;
        }
        if (at >= file.size() || !isdigit(file[at])) return false;
        v = 0;
        while (at < file.size() && isdigit(file[at])) v = v * 10 + (file[at++] - '0');
        return true;
    };
    long w, h, maxv;
    if (!next_int(w) || !next_int(h) || !next_int(maxv) || w <= 0 || h <= 0 || w > 65535 || h > 65535 || maxv <= 0 || maxv > 65535) {
        err = "bad PNM header";
        return false;
    }
    ++at;  // single whitespace after maxval
    const int nch = file[1] == '6' ? 3 : 1, bytes = maxv > 255 ? 2 : 1;
    const size_t need = (size_t)w * h * nch * bytes;
    if (at + need > file.size()) { err = "truncated PNM data"; return false; }
    img.w = (int)w; img.h = (int)h; img.channels = nch; img.bits = bytes * 8;
    img.px.resize((size_t)w * h * nch);
    const uint8_t* p = &file[at];
    for (size_t i = 0; i < img.px.size(); ++i) img.px[i] = bytes == 2 ? (uint16_t)(p[2 * i] << 8 | p[2 * i + 1]) : p[i];
    return true;
}

bool load_image(const std::string& path, Image& img, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { err = "cannot open"; return false; }
    if (file.size() >= 2 && file[0] == 'P') return decode_pnm(file, img, err);
    return decode_png(file, img, err);
}

// cv::imread(path) default flag: 3-channel 8-bit BGR whatever the file holds (HFTest.cpp:1241).
void to_bgr8(const Image& img, uint8_t* bgr) {
    const size_t n = (size_t)img.w * img.h;
    const int sh = img.bits == 16 ? 8 : 0;
    for (size_t i = 0; i < n; ++i) {
        const uint16_t* s = &img.px[i * img.channels];
        uint8_t r, g, b;
        if (img.channels >= 3) { r = (uint8_t)(s[0] >> sh); g = (uint8_t)(s[1] >> sh); b = (uint8_t)(s[2] >> sh); }
        else r = g = b = (uint8_t)(s[0] >> sh);
        bgr[i * 3 + 0] = b; bgr[i * 3 + 1] = g; bgr[i * 3 + 2] = r;
    }
}

// cv::imread(path, ANYDEPTH | ANYCOLOR) of a single-channel 16-bit depth image, read as ushort millimetres
// (HFTest.cpp:1247, :376).  Colour depth files have no defined meaning in the reference (it reads at<ushort>).
bool to_depth16(const Image& img, uint16_t* depth, std::string& err) {
    if (img.channels != 1) { err = "depth image must have one channel"; return false; }
    memcpy(depth, img.px.data(), (size_t)img.w * img.h * 2);
    return true;
}

bool write_png_rgb(const std::string& path, const uint8_t* bgr, int w, int h) {
    std::vector<uint8_t> raw(((size_t)w * 3 + 1) * h);
    for (int y = 0; y < h; ++y) {
        uint8_t* line = &raw[((size_t)w * 3 + 1) * y];
        line[0] = 0;
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = &bgr[((size_t)y * w + x) * 3];
            line[1 + x * 3] = s[2]; line[2 + x * 3] = s[1]; line[3 + x * 3] = s[0];
        }
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 1) != Z_OK) return false;
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    auto chunk = [&](const char* type, const uint8_t* body, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, body, len);
        const uint8_t tail[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        f.write(reinterpret_cast<const char*>(hdr), 8);
        if (len) f.write(reinterpret_cast<const char*>(body), len);
        f.write(reinterpret_cast<const char*>(tail), 4);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write(reinterpret_cast<const char*>(sig), 8);
    const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                              (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
    return (bool)f;
}

// 16-bit grey PNG (the renderer's depth<N>.png: millimetres, cv::imwrite of a CV_16U image)
bool write_png_gray16(const std::string& path, const uint16_t* px, int w, int h) {
    std::vector<uint8_t> raw(((size_t)w * 2 + 1) * h);
    for (int y = 0; y < h; ++y) {
        uint8_t* line = &raw[((size_t)w * 2 + 1) * y];
        line[0] = 0;
        for (int x = 0; x < w; ++x) {
            const uint16_t v = px[(size_t)y * w + x];
            line[1 + 2 * x] = (uint8_t)(v >> 8);
            line[2 + 2 * x] = (uint8_t)v;
        }
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 1) != Z_OK) return false;
    std::ofstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    auto chunk = [&](const char* type, const uint8_t* body, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, body, len);
        const uint8_t tail[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        f.write(reinterpret_cast<const char*>(hdr), 8);
        if (len) f.write(reinterpret_cast<const char*>(body), len);
        f.write(reinterpret_cast<const char*>(tail), 4);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write(reinterpret_cast<const char*>(sig), 8);
    const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                              (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 16, 0, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
    return (bool)f;
}

// ------------------------------------------------------------------------------------------------ text output
// Eigen's `operator<<` for a Matrix4f with the default IOFormat: stream precision (6 significant digits), every
// coefficient right-aligned to the widest one, one space between columns, '\n' between rows (HFTest.cpp:1286).
std::string eigen_format(const float m[16]) {
    std::string cell[16];
    size_t width = 0;
    for (int i = 0; i < 16; ++i) {
        std::ostringstream s;
        s << m[i];
        cell[i] = s.str();
        width = std::max(width, cell[i].size());
    }
    std::string out;
    for (int r = 0; r < 4; ++r) {
        if (r) out += '\n';
        for (int c = 0; c < 4; ++c) {
            if (c) out += ' ';
            out.append(width - cell[r * 4 + c].size(), ' ');
            out += cell[r * 4 + c];
        }
    }
    return out;
}

// GetOutName, HFTest.cpp:1135-1143
std::string out_stem(const std::string& filename) {
    std::string res = filename;
    size_t p = filename.find_last_of('/');
    if (p != std::string::npos) res = filename.substr(p + 1);
    p = res.find_last_of('.');
    if (p != std::string::npos) res = res.substr(0, p);
    return res;
}

// ------------------------------------------------------------------------------------------------ flags
struct Flags {
    bool test = false, train = false, other_mode = false, help = false, stage_times = false, check_inputs = false;
    std::string options_file, output_folder;
    int device = -1, encoder_mode = 0, feature_storage = -1;  // -1: the library's default
    // --train (main.cpp:13-25)
    std::string input, output = ".";
    int trees = 3, min_samples = 30, tests_per_node = 30, thresholds_per_test = 10, start_tree_no = 0, patch_size_in_voxels = -1;
    double voxel_size_in_m = -1;
    unsigned long long seed = 1;
    // --render (PatchGen/src/main.cpp:15-49)
    bool render = false, above_z = false, below_z = false, render_around_0 = false;
    int tessel_level = 1, in_place_rot = 24, lightings = 3, num_heights = 4;
    double height_step = 0.25, start_height = 0.3, object_radius = -1.0;
    // --genpatches / --gentrainpatches (PatchGen/src/main.cpp:16-37)
    bool genpatches = false, gentrainpatches = false, lmdb = false, binfile = false, no_random_values = false, use_surface_normals = false;
    int patch_size = 20, stride = 10, gpu = -1, batch_size = 1;
    double voxel_size = 0.001, distance_threshold = 3.0, max_depth_range_in_m = 0.25, percent = 1.0;
    std::string caffe_definition, caffe_weights;
};

bool parse_bool(const std::string& v) { return !(v == "false" || v == "0" || v == "no" || v == "f" || v == "n"); }

// gflags syntax: -flag / --flag, --flag=value or --flag value, --noflag for booleans.
bool parse_flags(int argc, char** argv, Flags& fl, std::string& err) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-') { err = "unexpected argument: " + a; return false; }
        a = a.substr(a[1] == '-' ? 2 : 1);
        std::string val;
        bool has_val = false;
        const size_t eq = a.find('=');
        if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        auto need = [&](std::string& dst) {
            if (!has_val) {
                if (i + 1 >= argc) { err = "flag --" + a + " needs a value"; return false; }
                val = argv[++i];
            }
            dst = val;
            return true;
        };
        std::string tmp;
        if (a == "test") fl.test = has_val ? parse_bool(val) : true;
        else if (a == "notest") fl.test = false;
        else if (a == "train") fl.train = has_val ? parse_bool(val) : true;
        else if (a == "learn_transitions" || a == "save_forest_map") fl.other_mode = true;
        else if (a == "detector_options_file" || a == "object_options_file") { if (!need(fl.options_file)) return false; }
        else if (a == "output_folder" || a == "output_dir") { if (!need(fl.output_folder)) return false; }  // README spelling
        else if (a == "device") { if (!need(tmp)) return false; fl.device = atoi(tmp.c_str()); }
        else if (a == "stage_times") fl.stage_times = has_val ? parse_bool(val) : true;
        else if (a == "encoder_mode") { if (!need(tmp)) return false; fl.encoder_mode = atoi(tmp.c_str()); }
        else if (a == "feature_storage") { if (!need(tmp)) return false; fl.feature_storage = atoi(tmp.c_str()); }
        else if (a == "check_inputs") fl.check_inputs = has_val ? parse_bool(val) : true;
        else if (a == "show_scene" || a == "visualize_hypotheses" || a == "noshow_scene" || a == "novisualize_hypotheses") {}
        else if (a == "help" || a == "h") fl.help = true;
        else if (a == "input") { if (!need(fl.input)) return false; }
        else if (a == "output") { if (!need(fl.output)) return false; }
        else if (a == "trees") { if (!need(tmp)) return false; fl.trees = atoi(tmp.c_str()); }
        else if (a == "min_samples") { if (!need(tmp)) return false; fl.min_samples = atoi(tmp.c_str()); }
        else if (a == "tests_per_node") { if (!need(tmp)) return false; fl.tests_per_node = atoi(tmp.c_str()); }
        else if (a == "thresholds_per_test") { if (!need(tmp)) return false; fl.thresholds_per_test = atoi(tmp.c_str()); }
        else if (a == "start_tree_no") { if (!need(tmp)) return false; fl.start_tree_no = atoi(tmp.c_str()); }
        else if (a == "patch_size_in_voxels") { if (!need(tmp)) return false; fl.patch_size_in_voxels = atoi(tmp.c_str()); }
        else if (a == "voxel_size_in_m") { if (!need(tmp)) return false; fl.voxel_size_in_m = atof(tmp.c_str()); }
        else if (a == "seed") { if (!need(tmp)) return false; fl.seed = strtoull(tmp.c_str(), nullptr, 10); }
        else if (a == "render") fl.render = has_val ? parse_bool(val) : true;
        else if (a == "tessel_level") { if (!need(tmp)) return false; fl.tessel_level = atoi(tmp.c_str()); }
        else if (a == "inPlaceCamRot") { if (!need(tmp)) return false; fl.in_place_rot = atoi(tmp.c_str()); }
        else if (a == "lightings") { if (!need(tmp)) return false; fl.lightings = atoi(tmp.c_str()); }
        else if (a == "numHeights") { if (!need(tmp)) return false; fl.num_heights = atoi(tmp.c_str()); }
        else if (a == "heightStep") { if (!need(tmp)) return false; fl.height_step = atof(tmp.c_str()); }
        else if (a == "startHeight") { if (!need(tmp)) return false; fl.start_height = atof(tmp.c_str()); }
        else if (a == "object_radius") { if (!need(tmp)) return false; fl.object_radius = atof(tmp.c_str()); }
        else if (a == "above_z") fl.above_z = has_val ? parse_bool(val) : true;
        else if (a == "below_z") fl.below_z = has_val ? parse_bool(val) : true;
        else if (a == "render_around_0") fl.render_around_0 = has_val ? parse_bool(val) : true;
        else if (a == "genpatches") fl.genpatches = has_val ? parse_bool(val) : true;
        else if (a == "gentrainpatches") fl.gentrainpatches = has_val ? parse_bool(val) : true;
        else if (a == "lmdb") fl.lmdb = has_val ? parse_bool(val) : true;
        else if (a == "binfile") fl.binfile = has_val ? parse_bool(val) : true;
        else if (a == "no_random_values") fl.no_random_values = has_val ? parse_bool(val) : true;
        else if (a == "use_surface_normals") fl.use_surface_normals = has_val ? parse_bool(val) : true;
        else if (a == "patch_size") { if (!need(tmp)) return false; fl.patch_size = atoi(tmp.c_str()); }
        else if (a == "stride") { if (!need(tmp)) return false; fl.stride = atoi(tmp.c_str()); }
        else if (a == "gpu") { if (!need(tmp)) return false; fl.gpu = atoi(tmp.c_str()); }
        else if (a == "batch_size") { if (!need(tmp)) return false; fl.batch_size = atoi(tmp.c_str()); }
        else if (a == "voxel_size") { if (!need(tmp)) return false; fl.voxel_size = atof(tmp.c_str()); }
        else if (a == "distance_threshold") { if (!need(tmp)) return false; fl.distance_threshold = atof(tmp.c_str()); }
        else if (a == "max_depth_range_in_m") { if (!need(tmp)) return false; fl.max_depth_range_in_m = atof(tmp.c_str()); }
        else if (a == "percent") { if (!need(tmp)) return false; fl.percent = atof(tmp.c_str()); }
        else if (a == "caffe_definition") { if (!need(fl.caffe_definition)) return false; }
        else if (a == "caffe_weights") { if (!need(fl.caffe_weights)) return false; }
        else if (a == "threads_per_tree" || a == "threads_for_parallel_trees" || a == "logtostderr" || a == "v" ||
                 a == "minloglevel") { if (!need(tmp)) return false; }  // CPU threading / glog flags: accepted, unused
        else { err = "unknown command line flag '" + a + "'"; return false; }
    }
    return true;
}

const char* kUsage =
    "usage: HoughForest --test --detector_options_file=<options.txt> [--output_folder=<dir>] [--device=<n>] [--stage_times]\n"
    "                   [--encoder_mode=0|1|2] 0: bf16 tensor-core operands (default), 1: split bf16, ~fp32 (3x encoder time),\n"
    "                                          2: fp16 operands\n"
    "                   [--feature_storage=0|1] feature rows between encoder and forest: 0 fp32, 1 fp16 (default where supported)\n"
    "       `rgb_path depth_path` pairs are read from stdin until EOF; results go to <dir><stem>_res.txt / _res.png\n"
    "       HoughForest --check_inputs   decode the stdin pairs only and print size + FNV-1a checksum of each frame\n"
    "       HoughForest --train --input=<training vectors> --output=<forest dir> --patch_size_in_voxels=<n> --voxel_size_in_m=<m>\n"
    "                   [--trees=3] [--min_samples=30] [--tests_per_node=30] [--thresholds_per_test=10] [--start_tree_no=0]\n"
    "                   [--seed=1] [--device=<n>]   trains on the GPU; forest.txt + tree<N>.dat as the reference writes them\n"
    "       HoughForest --render --input=<mesh.ply> --output=<dir> [--tessel_level=1] [--inPlaceCamRot=24] [--lightings=3]\n"
    "                   [--numHeights=4] [--heightStep=0.25] [--startHeight=0.3] [--above_z] [--below_z] [--render_around_0]\n"
    "                   [--object_radius=<m>]   PatchGen --render on the GPU: rgb<N>.png depth<N>.png pose<N>.txt per view\n"
    "       HoughForest --genpatches --input=<view dir of object 0>,<of object 1>,.. --output=<dir> [--lmdb | --binfile]\n"
    "                   [--patch_size=20] [--voxel_size=0.001] [--stride=10] [--no_random_values] [--distance_threshold=3]\n"
    "                   [--max_depth_range_in_m=0.25] [--percent=1] [--seed=1]   PatchGen --genpatches: data.mdb (LMDB, Datum\n"
    "                   per patch) + patch_annotation_lmdb.txt + patch_info_lmdb.txt, patches extracted on the GPU\n"
    "       HoughForest --gentrainpatches --caffe_weights=<weights> --input=<lmdb dir> --output=<training vectors>\n"
    "                   [--batch_size=1] [--gpu=0] [--encoder_mode=0|1|2]   PatchGen --gentrainpatches, encoder on the GPU\n";

uint64_t fnv1a(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// Host-only: decode every stdin pair exactly as --test would hand it to hf6d_submit and report what was read.
int check_inputs() {
    std::string rgb_fname, depth_fname, err;
    int bad = 0;
    while (std::cin >> rgb_fname >> depth_fname) {
        Image rgb_img, depth_img;
        if (!load_image(rgb_fname, rgb_img, err)) { std::cout << "Cannot read file: " << rgb_fname << " (" << err << ")" << std::endl; ++bad; continue; }
        if (!load_image(depth_fname, depth_img, err)) { std::cout << "Cannot read file: " << depth_fname << " (" << err << ")" << std::endl; ++bad; continue; }
        std::vector<uint8_t> bgr((size_t)rgb_img.w * rgb_img.h * 3);
        std::vector<uint16_t> depth((size_t)depth_img.w * depth_img.h);
        to_bgr8(rgb_img, bgr.data());
        if (!to_depth16(depth_img, depth.data(), err)) { std::cout << "Cannot read file: " << depth_fname << " (" << err << ")" << std::endl; ++bad; continue; }
        char line[256];
        snprintf(line, sizeof line, "bgr %dx%d %016llx depth %dx%d %016llx", rgb_img.w, rgb_img.h,
                 (unsigned long long)fnv1a(bgr.data(), bgr.size()), depth_img.w, depth_img.h,
                 (unsigned long long)fnv1a(depth.data(), depth.size() * 2));
        std::cout << line << std::endl;
    }
    return bad ? 4 : 0;
}

struct Ranked {
    float final_score;
    int index;
};

// MeshUtils::renderObject with alpha = 1 (MeshUtils.cpp:503-538, called from HFTest.cpp:1299): every pixel that a mesh vertex
// projects onto in front of the scene (or where the scene has no depth) gets a saturated green channel.  With alpha = 1 the
// reference's per-point update `g = alpha + (1 - alpha) * G / 255; G = min(g * 255 + g, 255)` is 255 whatever the order.  The
// object's name (cv::putText) is not drawn.
void render_object(uint8_t* bgr, const uint16_t* depth, int W, int H, const hf6d_params& p, const std::vector<float>& xyz,
                   const float pose[16]) {
    for (size_t i = 0; i + 2 < xyz.size(); i += 3) {
        const float x = pose[0] * xyz[i] + pose[1] * xyz[i + 1] + pose[2] * xyz[i + 2] + pose[3];
        const float y = pose[4] * xyz[i] + pose[5] * xyz[i + 1] + pose[6] * xyz[i + 2] + pose[7];
        const float z = pose[8] * xyz[i] + pose[9] * xyz[i + 1] + pose[10] * xyz[i + 2] + pose[11];
        const float rowf = y * p.fy / z + p.cy, colf = x * p.fx / z + p.cx;
        if (!(std::fabs(rowf) < 1e9f) || !(std::fabs(colf) < 1e9f)) continue;
        const int row = (int)rowf, col = (int)colf;
        if (row < 0 || row >= H || col < 0 || col >= W) continue;
        const uint16_t d = depth[(size_t)row * W + col];
        if (d == 0 || z < (float)d / 1000.0f) bgr[((size_t)row * W + col) * 3 + 1] = 255;
    }
}

}  // namespace

int main(int argc, char** argv) {
    Flags fl;
    std::string err;
    if (!parse_flags(argc, argv, fl, err)) {
        std::cerr << "ERROR: " << err << "\n" << kUsage;
        return 1;
    }
    if (fl.help) { std::cout << kUsage; return 0; }
    if (fl.other_mode) {
        std::cerr << "HoughForest: --learn_transitions / --save_forest_map have no implementation in the reference either (main.cpp:11-12)\n";
        return 2;
    }
    if (fl.render) {  // PatchGen/src/main.cpp:62-81
        if (fl.input.empty() || fl.output.empty()) { std::cerr << "Check failed: --render needs --input=<mesh.ply> and --output=<dir>" << std::endl; return 1; }
        if (fl.tessel_level <= 0) { std::cerr << "Check failed: FLAGS_tessel_level > 0" << std::endl; return 1; }
        if (fl.in_place_rot <= 0) { std::cerr << "Check failed: FLAGS_inPlaceCamRot > 0" << std::endl; return 1; }
        hf6d_render_params rp;
        hf6d_default_render_params(&rp);
        rp.tesselation_level = fl.tessel_level; rp.in_place_rotations = fl.in_place_rot; rp.lightings = fl.lightings;
        rp.heights = fl.num_heights; rp.height_step = (float)fl.height_step; rp.start_height = (float)fl.start_height;
        rp.above_z = fl.above_z; rp.below_z = fl.below_z; rp.render_around_0 = fl.render_around_0;
        rp.object_radius = (float)fl.object_radius; rp.device = std::max(fl.device, 0);
        hf6d_renderer* rd = nullptr;
        if (hf6d_renderer_create_ply(&rp, fl.input.c_str(), &rd)) {
            std::cerr << "HoughForest: cannot render " << fl.input << ": " << hf6d_last_error(nullptr) << std::endl;
            return 3;
        }
        const int nv = hf6d_renderer_view_count(rd);
        std::cout << "Total number of viewpoints: " << (long long)nv * rp.lightings << std::endl;  // .cpp:239-240
        std::vector<uint8_t> img((size_t)rp.W * rp.H * 3);
        std::vector<uint16_t> dep((size_t)rp.W * rp.H);
        int fcounter = 0, rc2 = 0;
        for (int v = 0; v < nv && !rc2; ++v) {
            double pose[16];
            hf6d_renderer_view(rd, v, pose);
            for (int light = 0; light < rp.lightings; ++light, ++fcounter) {  // SetAmbient(light * 0.1), .cpp:315
                if (hf6d_render(rd, pose, (float)(light * 0.1), img.data(), dep.data())) {
                    std::cerr << "HoughForest: " << hf6d_last_error(nullptr) << std::endl;
                    rc2 = 3;
                    break;
                }
                const std::string n = std::to_string(fcounter);
                if (!write_png_rgb(fl.output + "/rgb" + n + ".png", img.data(), rp.W, rp.H) ||
                    !write_png_gray16(fl.output + "/depth" + n + ".png", dep.data(), rp.W, rp.H)) {
                    std::cerr << "Check failed: cannot write into " << fl.output << std::endl;
                    rc2 = 1;
                    break;
                }
                std::ofstream fpose((fl.output + "/pose" + n + ".txt").c_str());  // the view transform, .cpp:126-137
                for (int i = 0; i < 4; ++i) {
                    for (int j = 0; j < 4; ++j) { fpose << pose[4 * i + j]; if (j != 3) fpose << " "; }
                    fpose << std::endl;
                }
            }
        }
        hf6d_renderer_destroy(rd);
        if (!rc2) std::cout << "Rendered " << fcounter << " views into " << fl.output << std::endl;
        return rc2;
    }
    if (fl.gentrainpatches) {  // PatchGen/src/main.cpp:117-136
        if (fl.caffe_weights.empty()) { std::cerr << "Check failed: FLAGS_caffe_weights.size() > 0 No caffe weights model defined." << std::endl; return 1; }
        if (fl.input.empty()) { std::cerr << "Check failed: FLAGS_input.size() > 0 No input lmdb specified." << std::endl; return 1; }
        if (fl.output.empty() || fl.output == ".") { std::cerr << "Check failed: FLAGS_output.size() > 0 No output file specified." << std::endl; return 1; }
        // --caffe_definition is accepted and unused: the layer shapes are in the weights file (encode1..3)
        hf6d_trainvec_stats st;
        const int dev = fl.gpu >= 0 ? fl.gpu : std::max(fl.device, 0);  // the reference's --gpu=-1 means Caffe on the CPU; there is no CPU path here
        if (hf6d_generate_train_vectors(fl.caffe_weights.c_str(), fl.input.c_str(), fl.output.c_str(), fl.batch_size, dev, fl.encoder_mode, &st)) {
            std::cerr << "HoughForest: " << hf6d_last_error(nullptr) << std::endl;
            return 3;
        }
        std::cout << "Total patches: " << st.written << std::endl;  // train_patch_generator.cpp:197
        return 0;
    }
    if (fl.genpatches) {  // PatchGen/src/main.cpp:83-115, patch_generator::generatePatches_rgbd
        std::vector<std::string> folders;
        {
            std::string cur;
            for (char ch : fl.input + ",") {  // boost::split(.., is_any_of(", "))
                if (ch == ',' || ch == ' ') { if (!cur.empty()) folders.push_back(cur); cur.clear(); }
                else cur += ch;
            }
        }
        if (folders.empty()) { std::cerr << "Check failed: input_object_folders_.size() != 0 No input objects specified" << std::endl; return 1; }
        if (fl.use_surface_normals) {
            std::cerr << "HoughForest: --use_surface_normals reads surface_normals<N>.bin files no renderer writes (the reference keeps the flag 'always set to false', PatchGen/src/main.cpp:41-42)" << std::endl;
            return 2;
        }
        if (!fl.lmdb && !fl.binfile) { std::cout << "No output method specified, using lmdb by default" << std::endl; fl.lmdb = true; }
        if (!fl.lmdb && !fl.no_random_values) {
            std::cerr << "HoughForest: --binfile holds the float patches of the texture gather, which has no random fill: add --no_random_values" << std::endl;
            return 2;
        }
        const std::string kind = fl.lmdb ? "lmdb" : "bin";
        std::ofstream finfo((fl.output + "/patch_info_" + kind + ".txt").c_str());
        if (!finfo) { std::cerr << "Check failed: finfo Cannot open file " << fl.output << "/patch_info_" << kind << ".txt for writing." << std::endl; return 1; }
        std::ofstream fannot((fl.output + "/patch_annotation_" + kind + ".txt").c_str());
        if (!fannot) { std::cerr << "Check failed: fannot Cannot open file " << fl.output << "/patch_annotation_" << kind << ".txt for writing." << std::endl; return 1; }
        fannot << folders.size() << std::endl;
        hf6d_patchdb* db = nullptr;
        if (fl.lmdb && hf6d_patchdb_create(fl.output.c_str(), &db)) {
            std::cerr << "Check failed: mdb_open failed. Does the lmdb already exist? (" << hf6d_last_error(nullptr) << ")" << std::endl;
            return 1;
        }
        const float voxel = (float)fl.voxel_size, range = (float)fl.max_depth_range_in_m;
        finfo << "Patch size in voxels: " << fl.patch_size << std::endl;
        finfo << "Voxel size in m: " << voxel << std::endl;
        finfo << "Stride in pixels: " << fl.stride << std::endl;
        finfo << "Max Depth Range: " << range << std::endl;
        hf6d_ctx* ex = nullptr;
        int W = 0, H = 0, rc2 = 0;
        const int ps = fl.patch_size, n_in = 4 * ps * ps;
        std::vector<uint8_t> bgr, q;
        std::vector<uint16_t> depth;
        std::vector<int32_t> locs;
        std::vector<float> fpatches;
        long long total = 0;
        auto keep = [&](int obj, int file, int i) {  // `rand() % 100 <= percent * 100` (patch_extractor.cu:378) with counted draws
            if (fl.percent >= 1.0) return true;
            uint64_t x = fl.seed * 0x9E3779B97F4A7C15ull + ((uint64_t)obj << 48 ^ (uint64_t)file << 24 ^ (uint64_t)i);
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
            return (double)(x % 100) <= fl.percent * 100;
        };
        for (size_t obj = 0; obj < folders.size() && !rc2; ++obj) {
            const std::string name = folders[obj].substr(folders[obj].find_last_of('/') + 1);
            std::cout << "Creating patches for object: " << name << std::endl;
            std::ofstream fbin;
            if (!fl.lmdb) {
                fbin.open((fl.output + "/" + name + ".bin").c_str(), std::ios::out | std::ios::binary);
                if (!fbin) { std::cerr << "Check failed: fout Output file " << fl.output << "/" << name << " cannot be openned." << std::endl; rc2 = 1; break; }
                fbin.write((const char*)&ps, 4); fbin.write((const char*)&voxel, 4); fbin.write((const char*)&range, 4);
            }
            long long n_obj = 0;
            int patch_id = 0;
            for (int file = 0;; ++file) {
                const std::string base = folders[obj] + "/", n = std::to_string(file);
                Image rgb, dep;
                std::string err;
                if (!load_image(base + "rgb" + n + ".png", rgb, err)) break;  // cv::imread(..).empty(): the end of the views
                if (!load_image(base + "depth" + n + ".png", dep, err)) { std::cerr << "Check failed: !depth.empty() File " << base << "depth" << n << ".png not exist, while the rgb file does." << std::endl; rc2 = 1; break; }
                std::ifstream fpose((base + "pose" + n + ".txt").c_str());
                float pose[16];
                bool pose_ok = (bool)fpose;
                for (int k = 0; k < 16 && pose_ok; ++k) pose_ok = (bool)(fpose >> pose[k]);
                if (!pose_ok) { std::cerr << "Check failed: fpose File " << base << "pose" << n << ".txt not exist, while the rgb and depth files does." << std::endl; rc2 = 1; break; }
                if (!ex) {
                    W = rgb.w; H = rgb.h;
                    hf6d_params p;
                    hf6d_default_params(&p);
                    p.W = W; p.H = H; p.stride = fl.stride;
                    p.fx = p.fy = (float)H / 2.0f / (float)tan((double)(45.3105f / 180.0f * 3.141592f / 2.0f));  // getFocalLength, patch_generator.h:43-45
                    p.cx = (float)W / 2.0f - 0.5f; p.cy = (float)H / 2.0f - 0.5f;
                    p.patch_vox = ps; p.voxel_m = voxel; p.max_depth_range_m = range; p.distance_threshold_m = (float)fl.distance_threshold;
                    p.fill_random = !fl.no_random_values; p.fill_seed = fl.seed; p.batch_size = 1; p.patch_mode = 0;
                    if (hf6d_create_extractor(&p, nullptr, std::max(fl.device, 0), &ex)) { std::cerr << "HoughForest: " << hf6d_last_error(nullptr) << std::endl; rc2 = 3; break; }
                    bgr.resize((size_t)W * H * 3); depth.resize((size_t)W * H);
                    const size_t cap = (size_t)hf6d_patch_capacity(ex);
                    q.resize(cap * n_in); locs.resize(cap * 2);
                }
                if (rgb.w != W || rgb.h != H || dep.w != W || dep.h != H) { std::cerr << "HoughForest: " << base << "rgb" << n << ".png is not " << W << " x " << H << " like the first view" << std::endl; rc2 = 1; break; }
                to_bgr8(rgb, bgr.data());
                if (!to_depth16(dep, depth.data(), err)) { std::cerr << "HoughForest: " << base << "depth" << n << ".png: " << err << std::endl; rc2 = 1; break; }
                int counts[2] = {0, 0};
                if (hf6d_upload(ex, 0, bgr.data(), depth.data()) || hf6d_run(ex, 0, HF6D_STAGE_SCAN, HF6D_STAGE_GATHER) ||
                    hf6d_fetch(ex, 0, HF6D_BUF_COUNTS, counts, sizeof counts) < 0 ||
                    hf6d_fetch(ex, 0, HF6D_BUF_LOCS, locs.data(), locs.size() * 4) < 0 ||
                    hf6d_fetch(ex, 0, HF6D_BUF_PATCH_U8, q.data(), q.size()) < 0) {
                    std::cerr << "HoughForest: " << hf6d_last_error(ex) << std::endl; rc2 = 3; break;
                }
                const int P = counts[1];
                if (!fl.lmdb) {
                    fpatches.resize((size_t)P * n_in);
                    if (P && hf6d_debug_texture_gather(ex, 0, fpatches.data(), fpatches.size() * 4) < 0) { std::cerr << "HoughForest: " << hf6d_last_error(ex) << std::endl; rc2 = 3; break; }
                }
                int kept = 0;
                for (int i = 0; i < P && !rc2; ++i) {
                    if (!keep((int)obj, file, i)) continue;
                    ++kept;
                    if (!fl.lmdb) { fbin.write((const char*)&fpatches[(size_t)i * n_in], (std::streamsize)n_in * 4); continue; }
                    char key[20];
                    snprintf(key, sizeof key, "%04d_%08d", (int)obj, patch_id++);
                    if (hf6d_patchdb_put(db, key, 4, ps, ps, (int)obj, &q[(size_t)i * n_in])) { std::cerr << "Check failed: mdb_put failed (" << hf6d_last_error(nullptr) << ")" << std::endl; rc2 = 1; break; }
                    const int x = locs[2 * i], y = locs[2 * i + 1];
                    float a[6];
                    hf6d_patch_annotation(W, H, 45.3105f, x, y, depth[(size_t)y * W + x], pose, a);
                    fannot << key << " " << a[0] << " " << a[1] << " " << a[2] << " " << a[3] << " " << a[4] << " " << a[5] << std::endl;
                }
                n_obj += kept;
                std::cout << "Extracted patches from file: " << file + 1 << "\r";
            }
            total += n_obj;
            std::cout << std::endl;
            std::cout << "Patches for " << name << ": " << n_obj << std::endl;
            finfo << "Patches for " << name << ": " << n_obj << std::endl;
        }
        if (ex) hf6d_destroy(ex);
        if (db && hf6d_patchdb_close(db) && !rc2) { std::cerr << "Check failed: mdb_txn_commit failed (" << hf6d_last_error(nullptr) << ")" << std::endl; rc2 = 1; }
        if (!rc2) {
            std::cout << "Finished! Total patches: " << total << std::endl;
            finfo << "Total patches: " << total << std::endl;
        }
        return rc2;
    }
    if (fl.train) {  // main.cpp:41-66
        const char* bad = fl.input.empty() ? "No input file specified"
                          : fl.output.empty() ? "No output folder specified"
                          : fl.trees <= 0 ? "You should train at least one tree"
                          : fl.min_samples <= 0 ? "min_samples should be greater than zero"
                          : fl.tests_per_node <= 0 ? "There should be at least one test per node"
                          : fl.thresholds_per_test <= 0 ? "There should be at least one threshold per test"
                          : fl.patch_size_in_voxels <= 0 ? "You should specify the Patch Size in Voxels (--patch_size_in_voxels)"
                          : fl.voxel_size_in_m <= 0 ? "Shoud should specify the Voxel Size in Meters (--voxel_size_in_m)" : nullptr;
        if (bad) { std::cerr << "Check failed: " << bad << std::endl; return 1; }
        hf6d_train_params tp;
        hf6d_default_train_params(&tp);
        tp.trees = fl.trees; tp.min_samples = fl.min_samples; tp.tests_per_node = fl.tests_per_node;
        tp.thresholds_per_test = fl.thresholds_per_test; tp.start_tree_no = fl.start_tree_no;
        tp.patch_size_in_voxels = fl.patch_size_in_voxels; tp.voxel_size_in_m = (float)fl.voxel_size_in_m;
        tp.seed = fl.seed; tp.device = std::max(fl.device, 0);
        std::cout << "Thread 0: Reading input..." << std::endl;  // HFTrain.cpp:1214
        hf6d_train_stats st;
        if (hf6d_train_forest(&tp, fl.input.c_str(), fl.output.c_str(), &st)) {
            std::cerr << "HoughForest: training failed: " << hf6d_last_error(nullptr) << std::endl;
            return 3;
        }
        for (int t = tp.start_tree_no; t < tp.start_tree_no + tp.trees; ++t) std::cout << "Tree " << t << " saved" << std::endl;
        std::cout << tp.trees << " trees, " << st.training_samples << " training samples each, " << st.nodes << " nodes, " << st.leaves
                  << " leaves, depth " << st.max_depth << ", " << st.train_ms / 1000.0 << "sec on the GPU" << std::endl;
        return 0;
    }
    if (fl.check_inputs) return check_inputs();
    if (!fl.test) return 0;  // main.cpp:70-76: nothing to do without a mode flag
    if (fl.options_file.empty()) {
        std::cerr << "Check failed: No detector options file specified (--detector_options_file)\n";
        return 1;
    }
    // HFTest.cpp:1190-1192: a '/' is appended to a non-empty output folder
    if (!fl.output_folder.empty() && fl.output_folder[fl.output_folder.size() - 1] != '/') fl.output_folder += '/';

    // model artefacts are validated on the host before a device is touched
    hf6d_options opt;
    std::vector<hf6d_object> objects(HF6D_MAX_CLASSES);
    if (hf6d_parse_options(fl.options_file.c_str(), &opt, objects.data(), (int)objects.size())) {
        std::cerr << "Cannot use options file " << fl.options_file << ": " << hf6d_last_error(nullptr) << std::endl;
        return 1;
    }
    objects.resize((size_t)std::min<int>(opt.n_objects, HF6D_MAX_CLASSES));
    hf6d_model_info forest;
    if (hf6d_inspect_forest(opt.forest_folder, &forest)) {
        std::cerr << "Cannot load forest: " << hf6d_last_error(nullptr) << std::endl;
        return 1;
    }
    int32_t dims[4];
    if (hf6d_inspect_weights(opt.caffe_weights, dims)) {
        std::cerr << "Cannot load encoder weights: " << hf6d_last_error(nullptr) << std::endl;
        return 1;
    }
    if (opt.n_objects != forest.K) {  // HFTest.cpp:1186
        std::cerr << "Check failed: Number of objects provided in the options file (" << opt.n_objects
                  << ") does not match the number of classes in the forest (" << forest.K << ")" << std::endl;
        return 1;
    }

    hf6d_ctx* ctx = nullptr;
    int ctx_w = 0, ctx_h = 0;
    uint8_t* bgr = nullptr;
    uint16_t* depth = nullptr;
    std::vector<hf6d_hypothesis> hyp(HF6D_MAX_CLASSES * HF6D_MAX_CENTRES * HF6D_MAX_HYPOTHESES_PER_CENTRE);
    std::vector<hf6d_detection> det(hyp.size());
    std::vector<std::vector<float>> vertices;  // per object: the mesh vertices (MeshUtils::renderObject projects them)
    bool refined = false;
    int rc = 0;

    std::string rgb_fname, depth_fname;
    while (std::cin >> rgb_fname >> depth_fname) {
        Image rgb_img, depth_img;
        if (!load_image(rgb_fname, rgb_img, err)) {
            std::cout << "Cannot read file: " << rgb_fname << std::endl;
            continue;
        }
        if (!load_image(depth_fname, depth_img, err)) {
            std::cout << "Cannot read file: " << depth_fname << std::endl;
            continue;
        }
        if (depth_img.w != rgb_img.w || depth_img.h != rgb_img.h) {
            std::cout << "Cannot read file: " << depth_fname << " (size differs from " << rgb_fname << ")" << std::endl;
            continue;
        }
        if (!ctx || rgb_img.w != ctx_w || rgb_img.h != ctx_h) {
            if (ctx) { hf6d_destroy(ctx); ctx = nullptr; }
            if (bgr) hf6d_host_free(bgr);
            if (depth) hf6d_host_free(depth);
            ctx_w = rgb_img.w; ctx_h = rgb_img.h;
            bgr = static_cast<uint8_t*>(hf6d_host_alloc((size_t)ctx_w * ctx_h * 3));
            depth = static_cast<uint16_t*>(hf6d_host_alloc((size_t)ctx_w * ctx_h * 2));
            if (hf6d_create_from_options(fl.options_file.c_str(), ctx_w, ctx_h, fl.device, 1, &ctx) || !bgr || !depth) {
                std::cerr << "HoughForest: cannot create the detector: " << hf6d_last_error(nullptr) << std::endl;
                return 3;
            }
            if ((fl.encoder_mode && hf6d_set_encoder_mode(ctx, fl.encoder_mode)) ||
                (fl.feature_storage >= 0 && hf6d_set_feature_storage(ctx, fl.feature_storage))) {
                std::cerr << "HoughForest: " << hf6d_last_error(ctx) << std::endl;
                return 3;
            }
            // MeshUtils::insertObjectFromPLY for every detected object (HFTest.cpp:1227-1233).  With the meshes the frame goes
            // through ICP, hypothesis scoring and the joint optimisation exactly where the reference runs them
            // (HFTest.cpp:927-934, :990-994) and _res.txt holds the refined poses; without them (the reference aborts on a
            // missing mesh) the Hough hypotheses are written, and the output says so.
            refined = hf6d_load_option_models(ctx) == 0;
            if (!refined) {
                std::cout << "Note: " << hf6d_last_error(ctx) << " -- poses are the pre-ICP Hough hypotheses, ranked by "
                             "pose_score_coeff * pose score + location_score_coeff * location score" << std::endl;
            } else {
                vertices.assign((size_t)forest.K, std::vector<float>());
                for (int k = 0; k < forest.K; ++k) {
                    if (!objects[k].should_detect) continue;
                    const int64_t bytes = hf6d_refine_fetch(ctx, HF6D_RBUF_MODEL_VERTICES, k, nullptr, 0);
                    if (bytes > 0) {
                        vertices[k].resize((size_t)bytes / 4);
                        hf6d_refine_fetch(ctx, HF6D_RBUF_MODEL_VERTICES, k, vertices[k].data(), (size_t)bytes);
                    }
                }
            }
        }
        to_bgr8(rgb_img, bgr);
        if (!to_depth16(depth_img, depth, err)) {
            std::cout << "Cannot read file: " << depth_fname << " (" << err << ")" << std::endl;
            continue;
        }

        const auto t0 = std::chrono::steady_clock::now();
        int n = 0, ticket = -1;
        if (hf6d_submit(ctx, bgr, depth, &ticket) || hf6d_wait(ctx, ticket, hyp.data(), (int)hyp.size(), &n)) {
            std::cerr << "HoughForest: detection failed on " << rgb_fname << ": " << hf6d_last_error(ctx) << std::endl;
            rc = 3;
            break;
        }
        n = std::min<int>(n, (int)hyp.size());
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        int32_t counts[2] = {0, 0};
        hf6d_fetch(ctx, 0, HF6D_BUF_COUNTS, counts, sizeof counts);
        std::cout << "Number of patches: " << counts[0] << std::endl;  // HFTest.cpp:427
        std::vector<hf6d_centre_list> centre_lists((size_t)forest.K);
        hf6d_fetch(ctx, 0, HF6D_BUF_CENTRES, centre_lists.data(), centre_lists.size() * sizeof(hf6d_centre_list));
        for (int k = 0; k < forest.K; ++k) {
            if (!objects[k].should_detect) continue;
            std::cout << "Generating Hypotheses for class: " << objects[k].name << std::endl;  // HFTest.cpp:697
            // max_loc_h = min(max_location_hypotheses, number of centre maxima), HFTest.cpp:710-712
            std::cout << "max locations: " << centre_lists[k].n << std::endl;
        }
        std::cout << "Total execution time: " << secs << "sec" << std::endl;  // HFTest.cpp:986
        if (fl.stage_times) {
            float ms[HF6D_STAGE_COUNT];
            static const char* names[HF6D_STAGE_COUNT] = {"scan", "gather", "encode", "traverse", "vote", "centres", "pose"};
            if (!hf6d_stage_ms(ctx, 0, ms))
                for (int s = 0; s < HF6D_STAGE_COUNT; ++s) std::cout << "  stage " << names[s] << ": " << ms[s] << " ms" << std::endl;
        }

        const std::string stem = out_stem(rgb_fname);
        const std::string out_fname = fl.output_folder + stem + "_res.txt";
        if (refined) {
            // HFTest.cpp:927-934 + :990-994: ICP, evaluate_hypothesis, optimize_hypotheses; then DetectObjects' output loop
            int nd = 0;
            if (hf6d_refine(ctx, 0, hyp.data(), n, det.data(), (int)det.size(), &nd)) {
                std::cerr << "HoughForest: refinement failed on " << rgb_fname << ": " << hf6d_last_error(ctx) << std::endl;
                rc = 3;
                break;
            }
            if (fl.stage_times) {
                float ms[4];
                static const char* names[4] = {"scene", "icp", "score", "optimise"};
                if (!hf6d_refine_ms(ctx, ms))
                    for (int s2 = 0; s2 < 4; ++s2) std::cout << "  stage refine/" << names[s2] << ": " << ms[s2] << " ms" << std::endl;
            }
            std::vector<int> by_rank;
            for (int i = 0; i < nd; ++i)
                if (det[i].rank >= 0) {
                    if ((int)by_rank.size() <= det[i].rank) by_rank.resize((size_t)det[i].rank + 1, -1);
                    by_rank[det[i].rank] = i;
                }
            std::ofstream fout(out_fname.c_str());
            if (!fout) {
                std::cerr << "Check failed: Cannot write to output file " << out_fname << std::endl;
                rc = 1;
                break;
            }
            std::cout << "Writing info to: " << out_fname << std::endl;
            std::vector<int> hcounter((size_t)forest.K, 0);
            int total_found = 0;
            for (int i : by_rank) {
                if (i < 0) continue;
                const hf6d_detection& d = det[i];
                ++hcounter[d.cls];
                ++total_found;
                fout << objects[d.cls].name << "(" << hcounter[d.cls] << ")" << ": " << std::endl;
                fout << eigen_format(d.pose) << std::endl;
                fout << std::endl;
                render_object(bgr, depth, ctx_w, ctx_h, opt.params, vertices[d.cls], d.pose);
            }
            fout.close();
            const std::string rgb_out_fname = fl.output_folder + stem + "_res.png";
            std::cout << "Writing result image to: " << rgb_out_fname << std::endl;
            if (!write_png_rgb(rgb_out_fname, bgr, ctx_w, ctx_h)) std::cerr << "cannot write " << rgb_out_fname << std::endl;
            std::cout << "Detection finished. Total objects found: " << total_found << std::endl;
            continue;
        }

        // HFTest.cpp:1261: sort by final score, descending (stable here, so equal scores keep emission order)
        std::vector<Ranked> order((size_t)n);
        for (int i = 0; i < n; ++i) {
            const float pose_score = (hyp[i].yawpitch_score + hyp[i].roll_score) / 2.0f;
            order[i] = Ranked{pose_score * opt.pose_score_coeff + hyp[i].loc_score * opt.location_score_coeff, i};
        }
        std::stable_sort(order.begin(), order.end(), [](const Ranked& a, const Ranked& b) { return a.final_score > b.final_score; });

        std::ofstream fout(out_fname.c_str());
        if (!fout) {
            std::cerr << "Check failed: Cannot write to output file " << out_fname << std::endl;
            rc = 1;
            break;
        }
        std::cout << "Writing info to: " << out_fname << std::endl;
        std::vector<int> hcounter((size_t)forest.K, 0);
        int total_found = 0;
        for (const Ranked& r : order) {
            const hf6d_hypothesis& h = hyp[r.index];
            if (hcounter[h.cls] < objects[h.cls].instances) {
                ++hcounter[h.cls];
                ++total_found;
                fout << objects[h.cls].name << "(" << hcounter[h.cls] << ")" << ": " << std::endl;
                fout << eigen_format(h.pose) << std::endl;
                fout << std::endl;
            }
        }
        fout.close();
        const std::string rgb_out_fname = fl.output_folder + stem + "_res.png";
        std::cout << "Writing result image to: " << rgb_out_fname << std::endl;
        if (!write_png_rgb(rgb_out_fname, bgr, ctx_w, ctx_h)) std::cerr << "cannot write " << rgb_out_fname << std::endl;
        std::cout << "Detection finished. Total objects found: " << total_found << std::endl;
    }
    if (ctx) hf6d_destroy(ctx);
    if (bgr) hf6d_host_free(bgr);
    if (depth) hf6d_host_free(depth);
    return rc;
}
