// scene_io.cpp — host-side readers in front of the C-ABI (include/ptb_sceneio.h): .scn scenes, OBJ/MTL and OFF meshes,
// texture / environment images.  Restates what the reference's readers leave in memory (file:line cited per function);
// written from the file formats, not from the reference's parser code: the reference scans with fscanf patterns, this
// reader is line/token based and yields the same values on files `Raytracer::save_scene` writes.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/ptb_sceneio.h"
#include "ptb_keyframes.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

bool read_file(const char* path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}
bool file_exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}
std::string dir_with_slash(const std::string& p) {       // extractFilePathWithEndingSlash (utils.cpp:20-25)
    const size_t k = p.find_last_of("/\\");
    return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}
std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s;
}
void copy_str(char* dst, const std::string& s) {
    const size_t n = std::min<size_t>(s.size(), PTB_PATH_MAX - 1);
    memcpy(dst, s.data(), n);
    dst[n] = 0;
}

// ------------------------------------------------------------------------------------------------ images
// Decoders return a TOP-DOWN 8-bit RGB image, the layout stbi_load(..., 3) gives the reference (utils.cpp:106).
struct Image {
    int W = 0, H = 0;
    std::vector<uint8_t> rgb;
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint32_t le32(const uint8_t* p) { return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]; }
uint32_t le16(const uint8_t* p) { return ((uint32_t)p[1] << 8) | p[0]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// decoded sizes come straight from file headers: cap them so that a malformed file is an error, not a std::bad_alloc
static const uint64_t PTB_MAX_IMAGE_PIXELS = (uint64_t)1 << 28;   // 16384 x 16384

int decode_png(const std::vector<uint8_t>& d, Image& im) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 33 || memcmp(d.data(), sig, 8)) return fail(PTB_ERR_INVALID, "png: bad signature");
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool have_ihdr = false;
    while (pos + 12 <= d.size()) {
        const uint32_t len = be32(&d[pos]);
        const uint8_t* type = &d[pos + 4];
        if (pos + 12 + (size_t)len > d.size()) return fail(PTB_ERR_INVALID, "png: truncated chunk");
        const uint8_t* body = &d[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail(PTB_ERR_INVALID, "png: short IHDR");
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) return fail(PTB_ERR_INVALID, "png: bad IHDR");
    if (interlace) return fail(PTB_ERR_UNSUPPORTED, "png: interlaced files are not decoded");
    int chan;
    switch (ctype) {
    case 0: chan = 1; break;
    case 2: chan = 3; break;
    case 3: chan = 1; break;
    case 4: chan = 2; break;
    case 6: chan = 4; break;
    default: return fail(PTB_ERR_INVALID, "png: bad colour type");
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))) || (ctype == 3 && depth == 16))
        return fail(PTB_ERR_INVALID, "png: bad bit depth");
    if ((uint64_t)w * (uint64_t)h > PTB_MAX_IMAGE_PIXELS) return fail(PTB_ERR_UNSUPPORTED, "png: image too large");
    const size_t bits = (size_t)depth * chan, stride = ((size_t)w * bits + 7) / 8, bpp = std::max<size_t>(1, bits / 8);
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return fail(PTB_ERR_INVALID, "png: inflate failed");
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    im.W = (int)w; im.H = (int)h;
    im.rgb.resize((size_t)w * h * 3);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* src = &raw[(stride + 1) * (size_t)y];
        const int ft = src[0];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= bpp ? prev[x - bpp] : 0;
            int v = src[1 + x];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: return fail(PTB_ERR_INVALID, "png: bad filter");
            }
            cur[x] = (uint8_t)v;
        }
        uint8_t* dst = &im.rgb[(size_t)y * w * 3];
        for (uint32_t x = 0; x < w; x++) {
            uint8_t s[4] = {0, 0, 0, 0};
            if (depth == 8) for (int k = 0; k < chan; k++) s[k] = cur[(size_t)x * chan + k];
            else if (depth == 16) for (int k = 0; k < chan; k++) s[k] = cur[((size_t)x * chan + k) * 2];   // high byte, as stb's 16->8 conversion
            else {
                const int per = 8 / depth, sh = (per - 1 - (int)(x % per)) * depth;
                const int v = (cur[x / per] >> sh) & ((1 << depth) - 1);
                s[0] = (uint8_t)(ctype == 3 ? v : v * (255 / ((1 << depth) - 1)));
            }
            if (ctype == 3) {
                const size_t e = (size_t)s[0] * 3;
                if (e + 3 > plte.size()) return fail(PTB_ERR_INVALID, "png: palette index out of range");
                dst[x * 3] = plte[e]; dst[x * 3 + 1] = plte[e + 1]; dst[x * 3 + 2] = plte[e + 2];
            } else if (chan <= 2) { dst[x * 3] = dst[x * 3 + 1] = dst[x * 3 + 2] = s[0]; }
            else { dst[x * 3] = s[0]; dst[x * 3 + 1] = s[1]; dst[x * 3 + 2] = s[2]; }
        }
        prev.swap(cur);
    }
    return PTB_OK;
}

int decode_bmp(const std::vector<uint8_t>& d, Image& im) {
    if (d.size() < 26 || d[0] != 'B' || d[1] != 'M') return fail(PTB_ERR_INVALID, "bmp: bad signature");
    const uint32_t off = le32(&d[10]), hsz = le32(&d[14]);
    int32_t w, h;
    int bpp, comp = 0;
    if (hsz == 12) { w = (int32_t)le16(&d[18]); h = (int32_t)le16(&d[20]); bpp = (int)le16(&d[24]); }
    else {
        if (d.size() < 54) return fail(PTB_ERR_INVALID, "bmp: short header");
        w = (int32_t)le32(&d[18]); h = (int32_t)le32(&d[22]); bpp = (int)le16(&d[28]); comp = (int)le32(&d[30]);
    }
    const bool flip = h > 0;
    h = abs(h);
    if (w <= 0 || h <= 0) return fail(PTB_ERR_INVALID, "bmp: bad size");
    if (!(comp == 0 || (comp == 3 && bpp == 32))) return fail(PTB_ERR_UNSUPPORTED, "bmp: compressed files are not decoded");
    if (bpp != 8 && bpp != 24 && bpp != 32) return fail(PTB_ERR_UNSUPPORTED, "bmp: only 8, 24 and 32 bits per pixel");
    const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
    if ((size_t)off + stride * (size_t)h > d.size()) return fail(PTB_ERR_INVALID, "bmp: truncated");
    const int pal_entry = hsz == 12 ? 3 : 4;
    // the palette starts right after the info header, whose size comes from the file: every entry a pixel names must lie inside it
    if (bpp == 8 && (uint64_t)14 + hsz + (uint64_t)pal_entry > d.size()) return fail(PTB_ERR_INVALID, "bmp: palette outside the file");
    const size_t pal_entries = bpp == 8 ? (d.size() - 14 - hsz) / pal_entry : 0;
    if ((uint64_t)w * (uint64_t)h > PTB_MAX_IMAGE_PIXELS) return fail(PTB_ERR_UNSUPPORTED, "bmp: image too large");
    const uint8_t* pal = d.data() + 14 + (bpp == 8 ? hsz : 0);
    im.W = w; im.H = h;
    im.rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; y++) {
        const uint8_t* src = &d[off + stride * (size_t)y];
        uint8_t* dst = &im.rgb[(size_t)(flip ? h - 1 - y : y) * w * 3];
        for (int x = 0; x < w; x++) {
            if (bpp == 8) { if (src[x] >= pal_entries) return fail(PTB_ERR_INVALID, "bmp: palette index outside the file"); const uint8_t* e = pal + (size_t)src[x] * pal_entry; dst[3 * x] = e[2]; dst[3 * x + 1] = e[1]; dst[3 * x + 2] = e[0]; }
            else { const uint8_t* e = src + (size_t)x * (bpp / 8); dst[3 * x] = e[2]; dst[3 * x + 1] = e[1]; dst[3 * x + 2] = e[0]; }
        }
    }
    return PTB_OK;
}

int decode_tga(const std::vector<uint8_t>& d, Image& im) {
    if (d.size() < 18) return fail(PTB_ERR_INVALID, "tga: short header");
    const int idlen = d[0], cmap = d[1], type = d[2], w = (int)le16(&d[12]), h = (int)le16(&d[14]), bpp = d[16], desc = d[17];
    if (cmap != 0 || !(type == 2 || type == 3 || type == 10 || type == 11)) return fail(PTB_ERR_UNSUPPORTED, "tga: only true-colour / grey images");
    const bool grey = type == 3 || type == 11, rle = type >= 10;
    if (w <= 0 || h <= 0 || !((grey && bpp == 8) || (!grey && (bpp == 24 || bpp == 32)))) return fail(PTB_ERR_UNSUPPORTED, "tga: unsupported pixel depth");
    const int bytes = bpp / 8;
    size_t pos = 18 + (size_t)idlen;
    if (!rle && pos + (size_t)w * h * bytes > d.size()) return fail(PTB_ERR_INVALID, "tga: truncated");
    std::vector<uint8_t> px((size_t)w * h * bytes);
    if (!rle) {
        if (pos + px.size() > d.size()) return fail(PTB_ERR_INVALID, "tga: truncated");
        memcpy(px.data(), &d[pos], px.size());
    } else {
        size_t o = 0;
        while (o < px.size()) {
            if (pos >= d.size()) return fail(PTB_ERR_INVALID, "tga: truncated RLE");
            const int c = d[pos++], n = (c & 127) + 1;
            if (c & 128) {
                if (pos + bytes > d.size()) return fail(PTB_ERR_INVALID, "tga: truncated RLE");
                for (int k = 0; k < n && o < px.size(); k++, o += bytes) memcpy(&px[o], &d[pos], bytes);
                pos += bytes;
            } else {
                const size_t m = std::min((size_t)n * bytes, px.size() - o);
                if (pos + m > d.size()) return fail(PTB_ERR_INVALID, "tga: truncated RLE");
                memcpy(&px[o], &d[pos], m);
                pos += m; o += m;
            }
        }
    }
    const bool top_down = (desc >> 5) & 1;
    im.W = w; im.H = h;
    im.rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; y++) {
        const uint8_t* src = &px[(size_t)y * w * bytes];
        uint8_t* dst = &im.rgb[(size_t)(top_down ? y : h - 1 - y) * w * 3];
        for (int x = 0; x < w; x++) {
            if (grey) dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = src[x];
            else { dst[3 * x] = src[x * bytes + 2]; dst[3 * x + 1] = src[x * bytes + 1]; dst[3 * x + 2] = src[x * bytes]; }
        }
    }
    return PTB_OK;
}

int decode_pnm(const std::vector<uint8_t>& d, Image& im) {
    if (d.size() < 7 || d[0] != 'P' || (d[1] != '5' && d[1] != '6')) return fail(PTB_ERR_INVALID, "pnm: only P5 / P6");
    const int chan = d[1] == '6' ? 3 : 1;
    size_t pos = 2;
    int vals[3];
    for (int k = 0; k < 3; k++) {
        for (;;) {
            while (pos < d.size() && isspace(d[pos])) pos++;
            if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') pos++; continue; }
            break;
        }
        int v = 0, nd = 0;
        while (pos < d.size() && isdigit(d[pos])) { v = v * 10 + (d[pos++] - '0'); nd++; }
        if (!nd) return fail(PTB_ERR_INVALID, "pnm: bad header");
        vals[k] = v;
    }
    pos++;   // the single whitespace after maxval
    const int w = vals[0], h = vals[1];
    if (w <= 0 || h <= 0 || vals[2] <= 0 || vals[2] > 255) return fail(PTB_ERR_UNSUPPORTED, "pnm: maxval must be 1..255");
    if (pos + (size_t)w * h * chan > d.size()) return fail(PTB_ERR_INVALID, "pnm: truncated");
    im.W = w; im.H = h;
    im.rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int k = 0; k < 3; k++) im.rgb[i * 3 + k] = d[pos + i * chan + (chan == 3 ? k : 0)];
    return PTB_OK;
}

// load_image<T> (utils.cpp:98-170): decode to 8-bit RGB, then swap rows top<->bottom.
int load_image_flipped(const char* path, Image& im) {
    std::vector<uint8_t> d;
    if (!path || !read_file(path, d)) return fail(PTB_ERR_INVALID, std::string("cannot read image file '") + (path ? path : "") + "'");
    int rc;
    if (d.size() >= 8 && d[0] == 0x89 && d[1] == 'P') rc = decode_png(d, im);
    else if (d.size() >= 2 && d[0] == 'B' && d[1] == 'M') rc = decode_bmp(d, im);
    else if (d.size() >= 2 && d[0] == 'P' && (d[1] == '5' || d[1] == '6')) rc = decode_pnm(d, im);
    else if (d.size() >= 3 && d[0] == 0xff && d[1] == 0xd8) rc = fail(PTB_ERR_UNSUPPORTED, std::string("JPEG images are not decoded: '") + path + "'");
    else if (lower(path).size() > 4 && lower(path).substr(lower(path).size() - 4) == ".tga") rc = decode_tga(d, im);
    else rc = fail(PTB_ERR_UNSUPPORTED, std::string("unknown image format: '") + path + "'");
    if (rc) return rc;
    const size_t row = (size_t)im.W * 3;
    for (int i = 0; i < im.H / 2; i++) std::swap_ranges(&im.rgb[row * i], &im.rgb[row * i] + row, &im.rgb[row * (size_t)(im.H - 1 - i)]);
    return PTB_OK;
}

// Texture::loadColors (BRDF.h:393-404) / loadNormals (406-419)
int load_texture_values(const char* path, int kind, std::vector<float>& values, int& W, int& H) {
    Image im;
    int rc = load_image_flipped(path, im);
    if (rc) return rc;
    W = im.W; H = im.H;
    values.resize(im.rgb.size());
    if (kind == 0) {
        for (size_t i = 0; i < values.size(); i++) { float v = (float)im.rgb[i]; v /= 255.f; values[i] = powf(v, 2.2f); }
    } else {
        for (size_t i = 0; i < values.size() / 3; i++) {
            // Vector::normalize (Vector.h:95-101): divide by sqrt(norm2)
            const float x = (float)im.rgb[i * 3] - 128, y = (float)im.rgb[i * 3 + 1] - 128, z = (float)im.rgb[i * 3 + 2] - 128;
            const float n = sqrtf(x * x + y * y + z * z);
            values[i * 3] = x / n; values[i * 3 + 1] = y / n; values[i * 3 + 2] = z / n;
        }
    }
    return PTB_OK;
}

// ------------------------------------------------------------------------------------------------ mesh files
struct Slot {
    std::string file;            // "" = constant
    float mult[3] = {1, 1, 1};
};
Slot const_slot(float a, float b, float c) { Slot s; s.mult[0] = a; s.mult[1] = b; s.mult[2] = c; return s; }

}  // namespace

struct ptb_meshfile {
    std::vector<float> vertices, normals, uvs, vertex_colors;
    std::vector<int32_t> tri;
    std::map<std::string, int> group_names;
    bool has_materials = false;
    std::vector<Slot> slots[PTB_N_KINDS];   // per kind, per group
};

namespace {

// One "a", "a/b", "a/b/c" or "a//c" vertex reference; returns the characters consumed (0 = none).
// Indices are signed like the reference's `%u` into an int (strtoul accepts a sign).
int parse_ref(const char* s, int& v, int& t, int& n, int& form) {
    const char* p = s;
    while (*p == ' ' || *p == '\t') p++;
    char* e;
    long a = strtol(p, &e, 10);
    if (e == p) return 0;
    v = (int)a; form = 0; p = e;
    if (*p == '/') {
        if (p[1] == '/') {
            long c = strtol(p + 2, &e, 10);
            if (e == p + 2) return (int)(p - s);
            n = (int)c; form = 3; p = e;
        } else {
            long b = strtol(p + 1, &e, 10);
            if (e == p + 1) return (int)(p - s);
            t = (int)b; form = 1; p = e;
            if (*p == '/') {
                long c = strtol(p + 1, &e, 10);
                if (e != p + 1) { n = (int)c; form = 2; p = e; }
            }
        }
    }
    return (int)(p - s);
}

int resolve(int idx, size_t count) { return idx < 0 ? (int)count + idx : idx - 1; }   // TriangleMesh.cpp:333-341

// TriMesh::readOBJ geometry part (TriangleMesh.cpp:240-475)
int read_obj(const char* path, ptb_meshfile& m, std::string& mtl) {
    FILE* f = fopen(path, "r");
    if (!f) return fail(PTB_ERR_INVALID, std::string("cannot open mesh file '") + path + "'");
    int cur_group = -1;
    bool bad_ref = false;
    std::string line;
    char buf[4096];
    while (fgets(buf, sizeof buf, f)) {
        line = buf;
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        const char* l = line.c_str();
        if (l[0] == 'u' && l[1] == 's') {                       // usemtl <name>: groups numbered by first appearance
            const char* p = l + 6;
            while (*p == ' ') p++;
            const std::string name(p);
            auto it = m.group_names.find(name);
            if (it != m.group_names.end()) cur_group = it->second;
            else { cur_group = (int)m.group_names.size(); m.group_names[name] = cur_group; }
        } else if (l[0] == 'm' && l[1] == 't' && l[2] == 'l') { // mtllib <file>
            const char* p = l + 6;
            while (*p == ' ') p++;
            mtl = p;
        } else if (l[0] == 'v' && l[1] == ' ') {
            float v[3] = {0, 0, 0}, c[3];
            if (sscanf(l, "v %f %f %f %f %f %f", &v[0], &v[1], &v[2], &c[0], &c[1], &c[2]) == 6) {
                for (int k = 0; k < 3; k++) m.vertex_colors.push_back(std::min(1.f, std::max(0.f, c[k])));
            }
            m.vertices.insert(m.vertices.end(), v, v + 3);
        } else if (l[0] == 'v' && l[1] == 'n') {
            float v[3] = {0, 0, 0};
            sscanf(l, "vn %f %f %f", &v[0], &v[1], &v[2]);
            m.normals.insert(m.normals.end(), v, v + 3);
        } else if (l[0] == 'v' && l[1] == 't') {
            float v[2] = {0, 0};
            sscanf(l, "vt %f %f", &v[0], &v[1]);
            m.uvs.insert(m.uvs.end(), v, v + 2);
        } else if (l[0] == 'f') {
            // first three references form the first triangle, every further one fans (i0, previous, new) (390-458)
            const size_t nv = m.vertices.size() / 3, nt = m.uvs.size() / 2, nn = m.normals.size() / 3;
            const char* p = l + 1;
            int v[3], t[3], n[3], form0 = -1, got = 0;
            for (; got < 3; got++) {
                int form;
                const int used = parse_ref(p, v[got], t[got], n[got], form);
                if (!used) break;
                if (got == 0) form0 = form;
                p += used;
            }
            if (got < 3) continue;
            auto emit = [&](int a, int b, int c, int form) {
                int32_t r[10] = {resolve(v[a], nv), resolve(v[b], nv), resolve(v[c], nv), -1, -1, -1, -1, -1, -1, cur_group};
                if (form == 1 || form == 2) { r[3] = resolve(t[a], nt); r[4] = resolve(t[b], nt); r[5] = resolve(t[c], nt); }
                if (form == 2 || form == 3) { r[6] = resolve(n[a], nn); r[7] = resolve(n[b], nn); r[8] = resolve(n[c], nn); }
                // a relative reference that points before the start of its list (e.g. "f -5/-5/-5" after two vt lines) resolves below
                // zero: the reference indexes out of bounds there; this reader refuses the file
                for (int q = 0; q < 9; q++) if (r[q] < -1 || (q < 3 && r[q] < 0)) bad_ref = true;
                m.tri.insert(m.tri.end(), r, r + 10);
            };
            emit(0, 1, 2, form0);
            for (;;) {
                while (*p == ' ' || *p == '\t') p++;
                if (!*p) break;
                int form, v3 = 0, t3 = 0, n3 = 0;
                const int used = parse_ref(p, v3, t3, n3, form);
                if (!used) { p++; continue; }
                p += used;
                v[1] = v[2]; t[1] = t[2]; n[1] = n[2];
                v[2] = v3; t[2] = t3; n[2] = n3;
                emit(0, 1, 2, form);
            }
        }
    }
    fclose(f);
    if (bad_ref) return fail(PTB_ERR_INVALID, "obj: a face references an element before the start of its list");
    if (m.group_names.empty()) {                                  // 470-475
        for (size_t i = 9; i < m.tri.size(); i += 10) m.tri[i] = 0;
        m.group_names["Default"] = 0;
    }
    return PTB_OK;
}

// MTL part of readOBJ (TriangleMesh.cpp:478-565).  Tests are on the RAW line characters like the reference's
// (`line[0]=='K' && line[1]=='d'`): indented statements are not recognised there either.
void read_mtl(const std::string& obj_path, const std::string& mtl, ptb_meshfile& m) {
    m.has_materials = true;
    const size_t ng = m.group_names.size();
    for (size_t g = 0; g < ng; g++) {                             // 481-490
        m.slots[PTB_KIND_KD].push_back(const_slot(.5f, .5f, .5f));
        m.slots[PTB_KIND_KS].push_back(const_slot(0, 0, 0));
        m.slots[PTB_KIND_NE].push_back(const_slot(0, 0, 0));
        m.slots[PTB_KIND_NORMAL].push_back(const_slot(0, 0, 1));
        m.slots[PTB_KIND_ALPHA].push_back(const_slot(1, 1, 1));
        m.slots[PTB_KIND_REFR].push_back(const_slot(1.3f, 1.3f, 1.3f));
        m.slots[PTB_KIND_TRANSP].push_back(const_slot(1, 1, 1));
        m.slots[PTB_KIND_SUBSURF].push_back(const_slot(0, 0, 0));
    }
    const std::string dir = dir_with_slash(obj_path);
    FILE* f = fopen((dir + mtl).c_str(), "r");
    if (!f) return;
    char buf[4096];
    std::string grp;
    // `groupNames[grp]` in the reference is std::map::operator[]: a material the OBJ never used is INSERTED with id 0
    auto gid = [&](const std::string& name) -> size_t {
        auto it = m.group_names.find(name);
        if (it == m.group_names.end()) { m.group_names[name] = 0; return 0; }
        return (size_t)it->second;
    };
    auto rest = [](const char* l, size_t skip) {
        std::string s(strlen(l) > skip ? l + skip : "");
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        size_t b = 0;
        while (b < s.size() && (s[b] == ' ' || s[b] == '\t')) b++;
        return s.substr(b);
    };
    while (fgets(buf, sizeof buf, f)) {
        const char* l = buf;
        const size_t len = strlen(l);
        if (l[0] == 'n' && l[1] == 'e' && l[2] == 'w') { grp = rest(l, 7); continue; }
        if (len > 5 && l[0] == 'm' && l[4] == 'K' && l[5] == 'd') m.slots[PTB_KIND_KD][gid(grp)].file = dir + rest(l, 7);
        if (len > 5 && l[0] == 'm' && l[4] == 'K' && l[5] == 's') m.slots[PTB_KIND_KS][gid(grp)].file = dir + rest(l, 7);
        if (len > 5 && l[0] == 'm' && l[4] == 'B' && l[5] == 'u') m.slots[PTB_KIND_NORMAL][gid(grp)].file = dir + rest(l, 9);
        if (len > 4 && l[0] == 'm' && l[1] == 'a' && l[4] == 'd') m.slots[PTB_KIND_ALPHA][gid(grp)].file = dir + rest(l, 6);
        if (l[0] == 'K' && l[1] == 'd') { float* q = m.slots[PTB_KIND_KD][gid(grp)].mult; sscanf(l, "Kd %f %f %f", &q[0], &q[1], &q[2]); }
        if (l[0] == 'K' && l[1] == 's') { float* q = m.slots[PTB_KIND_KS][gid(grp)].mult; sscanf(l, "Ks %f %f %f", &q[0], &q[1], &q[2]); }
        if (l[0] == 'N' && l[1] == 's') {
            float q[3] = {0, 0, 0};
            const int r = sscanf(l, "Ns %f %f %f", &q[0], &q[1], &q[2]);
            if (r == 1) q[1] = q[2] = q[0];
            memcpy(m.slots[PTB_KIND_NE][gid(grp)].mult, q, sizeof q);
        }
    }
    fclose(f);
}

// TriMesh::readOFF (TriangleMesh.cpp:107-130): a token stream; every face consumes exactly four integers
int read_off(const char* path, ptb_meshfile& m) {
    FILE* f = fopen(path, "r");
    if (!f) return fail(PTB_ERR_INVALID, std::string("cannot open mesh file '") + path + "'");
    char tag[64];
    int nv = 0, nf = 0, nx = 0;
    if (fscanf(f, "%63s", tag) != 1 || fscanf(f, "%d %d %d", &nv, &nf, &nx) != 3 || nv < 0 || nf < 0) { fclose(f); return fail(PTB_ERR_INVALID, "off: bad header"); }
    {   // every vertex takes at least six characters of the file: a header that promises more than the file can hold is malformed
        const long here = ftell(f);
        fseek(f, 0, SEEK_END);
        const long size = ftell(f);
        fseek(f, here, SEEK_SET);
        if ((int64_t)nv * 6 > (int64_t)size || (int64_t)nf * 8 > (int64_t)size) { fclose(f); return fail(PTB_ERR_INVALID, "off: header counts exceed the file"); }
    }
    m.vertices.resize((size_t)nv * 3);
    for (int i = 0; i < nv * 3; i++) if (fscanf(f, "%f", &m.vertices[i]) != 1) { fclose(f); return fail(PTB_ERR_INVALID, "off: truncated vertices"); }
    for (int i = 0; i < nf; i++) {
        int k, a, b, c;
        if (fscanf(f, "%d %d %d %d", &k, &a, &b, &c) != 4) { fclose(f); return fail(PTB_ERR_INVALID, "off: truncated faces"); }
        const int32_t r[10] = {a, b, c, -1, -1, -1, -1, -1, -1, -1};
        m.tri.insert(m.tri.end(), r, r + 10);
    }
    fclose(f);
    return PTB_OK;
}

int read_meshfile(const char* path, int load_textures, ptb_meshfile& m) {
    const std::string lo = lower(path);
    if (lo.find(".off") != std::string::npos) return read_off(path, m);
    if (lo.find(".wrl") != std::string::npos) return fail(PTB_ERR_UNSUPPORTED, "VRML meshes are not read");
    if (lo.find(".obj") != std::string::npos) {
        std::string mtl;
        int rc = read_obj(path, m, mtl);
        if (rc) return rc;
        if (load_textures) read_mtl(path, mtl, m);
        return PTB_OK;
    }
    return fail(PTB_ERR_UNSUPPORTED, std::string("unknown mesh format: '") + path + "'");
}

// ------------------------------------------------------------------------------------------------ .scn
struct Keyframes {
    std::vector<std::pair<float, float>> scale;
    std::vector<std::pair<float, std::vector<float>>> translation, rotation;
};
struct ScnObject {
    ptb_scn_object o;
    std::vector<Slot> slots[PTB_N_KINDS];
    Keyframes keys;
};

struct Lines {
    std::vector<std::string> l;
    size_t pos = 0;
    bool next(std::string& out) {                 // fscanf's "\n" eats blank lines and leading blanks of the next one
        while (pos < l.size()) {
            std::string s = l[pos++];
            while (!s.empty() && (s.back() == '\r' || s.back() == '\n' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
            size_t b = 0;
            while (b < s.size() && (s[b] == ' ' || s[b] == '\t')) b++;
            if (b == s.size()) continue;
            out = s.substr(b);
            return true;
        }
        return false;
    }
    void unread() { if (pos) pos--; }
};

bool starts(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }
std::string after(const std::string& s, const char* key) {   // text after "key:" with blanks trimmed
    size_t k = strlen(key);
    while (k < s.size() && (s[k] == ' ' || s[k] == '\t')) k++;
    return s.substr(std::min(k, s.size()));
}

}  // namespace

struct ptb_scn {
    ptb_scn_header h;
    std::vector<ScnObject> objects;
    std::string dir;
};

namespace {

#define NEED(cond, what) do { if (!(cond)) return fail(PTB_ERR_INVALID, std::string("scn: expected ") + what + " near line " + std::to_string(L.pos)); } while (0)

// the three keyframe maps of an object as tracks (ptb_keyframes.h): 0 scale, 1 translation, 2 rotation
static void keys_to_tracks(const Keyframes& k, ptb::KeyTrack tr[3]) {
    std::vector<float> fr, val;
    for (auto& e : k.scale) { fr.push_back(e.first); val.push_back(e.second); }
    ptb::key_track_set(tr[0], fr.data(), val.data(), (int)fr.size(), 1);
    fr.clear(); val.clear();
    for (auto& e : k.translation) { fr.push_back(e.first); val.insert(val.end(), e.second.begin(), e.second.end()); }
    ptb::key_track_set(tr[1], fr.data(), val.data(), (int)fr.size(), 3);
    fr.clear(); val.clear();
    for (auto& e : k.rotation) { fr.push_back(e.first); val.insert(val.end(), e.second.begin(), e.second.end()); }
    ptb::key_track_set(tr[2], fr.data(), val.data(), (int)fr.size(), 9);
}

int parse_slots(Lines& L, std::string& s, ScnObject& so, int kind, const char* count_key, bool line_in_s) {
    unsigned n = 0;
    if (!line_in_s) NEED(L.next(s), count_key);
    NEED(sscanf(s.c_str(), (std::string(count_key) + " %u").c_str(), &n) == 1, count_key);
    for (unsigned i = 0; i < n; i++) {
        NEED(L.next(s) && starts(s, "texture:"), "texture:");
        const std::string val = after(s, "texture:");
        Slot sl;
        float c[3], one;
        switch (kind) {                                              // Geometry.h:572-661
        case PTB_KIND_KD: case PTB_KIND_KS: case PTB_KIND_SUBSURF:
            if (starts(val, "Co") && sscanf(val.c_str(), "Color: (%f, %f, %f)", &c[0], &c[1], &c[2]) == 3) sl = const_slot(c[0] / 255.f, c[1] / 255.f, c[2] / 255.f);
            else sl.file = val;
            break;
        case PTB_KIND_NE:
            if (starts(val, "Co") && sscanf(val.c_str(), "Color: (%f, %f, %f)", &c[0], &c[1], &c[2]) == 3) sl = const_slot(c[0], c[1], c[2]);
            else sl.file = val;
            break;
        case PTB_KIND_NORMAL:
            sl = const_slot(0, 0, 1);
            if (!starts(val, "Nu")) sl.file = val;
            break;
        case PTB_KIND_ALPHA:
            if (sscanf(val.c_str(), "%f", &one) == 1) sl = const_slot(one, one, one);
            else sl.file = val;
            break;
        default: sl.file = val; break;                               // transparency / refraction index maps
        }
        if (sl.file == "Null") sl.file.clear();                      // the name save_to_file writes for constant slots
        NEED(L.next(s) && starts(s, "multiplier:"), "multiplier:");
        if (kind == PTB_KIND_TRANSP || kind == PTB_KIND_REFR) NEED(sscanf(s.c_str(), "multiplier: %f", &sl.mult[0]) == 1, "multiplier value");
        else NEED(sscanf(s.c_str(), "multiplier: (%f, %f, %f)", &sl.mult[0], &sl.mult[1], &sl.mult[2]) == 3, "multiplier triple");
        so.slots[kind].push_back(sl);
    }
    return PTB_OK;
}

// Object::load_from_file (Geometry.h:518-662)
int parse_object_common(Lines& L, ScnObject& so, const char* replaced) {
    ptb_scn_object& o = so.o;
    std::string s;
    unsigned u = 0;
    NEED(L.next(s) && starts(s, "name:"), "name:");
    std::string name = after(s, "name:");
    if (replaced) { const size_t k = name.find('#'); if (k != std::string::npos) name.replace(k, 1, replaced); }
    copy_str(o.name, name);
    NEED(L.next(s) && sscanf(s.c_str(), "miroir: %u", &u) == 1, "miroir:"); o.miroir = (int)u;
    NEED(L.next(s), "ghost: / translation:");
    if (s[0] == 'g') { NEED(sscanf(s.c_str(), "ghost: %u", &u) == 1, "ghost:"); o.ghost = (int)u; NEED(L.next(s), "translation:"); }
    float* tr = o.xform.translation; float* r = o.xform.rotation; float* c = o.xform.rotation_center;
    NEED(sscanf(s.c_str(), "translation: (%f, %f, %f)", &tr[0], &tr[1], &tr[2]) == 3, "translation:");
    NEED(L.next(s) && sscanf(s.c_str(), "rotation: (%f, %f, %f, %f, %f, %f, %f, %f, %f)", &r[0], &r[1], &r[2], &r[3], &r[4], &r[5], &r[6], &r[7], &r[8]) == 9, "rotation:");
    NEED(L.next(s) && sscanf(s.c_str(), "center: (%f, %f, %f)", &c[0], &c[1], &c[2]) == 3, "center:");
    NEED(L.next(s) && sscanf(s.c_str(), "scale: %f", &o.xform.scale) == 1, "scale:");
    NEED(L.next(s) && sscanf(s.c_str(), "display_edges: %u", &u) == 1, "display_edges:"); o.display_edges = (int)u;
    NEED(L.next(s) && sscanf(s.c_str(), "interp_normals: %u", &u) == 1, "interp_normals:"); o.interp_normals = (int)u;
    NEED(L.next(s) && sscanf(s.c_str(), "flip_normals: %u", &u) == 1, "flip_normals:"); o.flip_normals = (int)u;
    NEED(L.next(s), "nb_transforms: / nb_textures:");
    if (s.size() > 4 && s[4] == 'r') {
        NEED(sscanf(s.c_str(), "nb_transforms: %u", &u) == 1, "nb_transforms:");
        o.n_keyframes = (int)u;
        for (unsigned i = 0; i < u; i++) { float fr, v; NEED(L.next(s) && sscanf(s.c_str(), "%f %f", &fr, &v) == 2, "scale keyframe"); so.keys.scale.push_back({fr, v}); }
        for (unsigned i = 0; i < u; i++) { float fr, v[3]; NEED(L.next(s) && sscanf(s.c_str(), "%f %f, %f, %f", &fr, &v[0], &v[1], &v[2]) == 4, "translation keyframe"); so.keys.translation.push_back({fr, std::vector<float>(v, v + 3)}); }
        for (unsigned i = 0; i < u; i++) {
            float fr, v[9];
            NEED(L.next(s) && sscanf(s.c_str(), "%f %f, %f, %f, %f, %f, %f, %f, %f, %f", &fr, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) == 10, "rotation keyframe");
            so.keys.rotation.push_back({fr, std::vector<float>(v, v + 9)});
        }
        NEED(L.next(s), "nb_textures:");
    }
    int rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_KD, "nb_textures:", true))) return rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_NORMAL, "nb_normalmaps:", false))) return rc;
    NEED(L.next(s), "nb_subsurfaces: / nb_specularmaps:");
    if (starts(s, "nb_subsurfaces:")) {
        if ((rc = parse_slots(L, s, so, PTB_KIND_SUBSURF, "nb_subsurfaces:", true))) return rc;
        NEED(L.next(s), "nb_specularmaps:");
    }
    if ((rc = parse_slots(L, s, so, PTB_KIND_KS, "nb_specularmaps:", true))) return rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_ALPHA, "nb_alphamaps:", false))) return rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_NE, "nb_expmaps:", false))) return rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_TRANSP, "nb_transpmaps:", false))) return rc;
    if ((rc = parse_slots(L, s, so, PTB_KIND_REFR, "nb_refrindexmaps:", false))) return rc;
    for (int k = 0; k < PTB_N_KINDS; k++) o.n_slots[k] = (int32_t)so.slots[k].size();
    // key-framed placement replaces the static one when keys exist (Geometry.h:258-312); ptb_scn_object reports it at frame 0
    ptb::KeyTrack tracks[3];
    keys_to_tracks(so.keys, tracks);
    ptb::key_eval(tracks[0], 0.f, &o.xform.scale);
    ptb::key_eval(tracks[1], 0.f, tr);
    ptb::key_eval(tracks[2], 0.f, r);
    return PTB_OK;
}

int parse_object(Lines& L, ScnObject& so, const char* replaced) {
    memset(&so.o, 0, sizeof(so.o));
    ptb_scn_object& o = so.o;
    o.interp_normals = 1; o.is_centered = 1; o.xform.scale = 1;
    std::string s;
    unsigned u = 0;
    NEED(L.next(s) && s.size() > 5, "NEW <object>");
    if (s[4] == 'M') o.type = PTB_SCN_MESH;                              // Geometry.cpp:11-28
    else if (s[4] == 'S') o.type = PTB_SCN_SPHERE;
    else if (s[4] == 'P' && s[5] == 'L') o.type = PTB_SCN_PLANE;
    else if (s[4] == 'P' && s[5] == 'O') o.type = PTB_SCN_POINTSET;
    else return fail(PTB_ERR_INVALID, "scn: unknown object header '" + s + "'");
    int rc = parse_object_common(L, so, replaced);
    if (rc) return rc;
    if (o.type == PTB_SCN_SPHERE) {                                      // Geometry.h:887-911
        NEED(L.next(s) && sscanf(s.c_str(), "is_envmap: %u", &u) == 1, "is_envmap:"); o.is_envmap = (int)u;
        NEED(L.next(s) && starts(s, "envmapfilename:"), "envmapfilename:");
        if (o.is_envmap) copy_str(o.envmap, after(s, "envmapfilename:"));
        NEED(L.next(s) && sscanf(s.c_str(), "O: (%f, %f, %f)", &o.O[0], &o.O[1], &o.O[2]) == 3, "O:");
        NEED(L.next(s) && sscanf(s.c_str(), "R: %f", &o.R) == 1, "R:");
        // Sphere::init (Geometry.h:856-873): the centre of rotation is the origin; an environment map flips the normals
        for (int k = 0; k < 3; k++) o.xform.rotation_center[k] = o.O[k];
        if (o.is_envmap) o.flip_normals = 1;
    } else if (o.type == PTB_SCN_PLANE) {                                // Geometry.h:1203-1213
        NEED(L.next(s) && sscanf(s.c_str(), "Point: (%f, %f, %f)", &o.A[0], &o.A[1], &o.A[2]) == 3, "Point:");
        NEED(L.next(s) && sscanf(s.c_str(), "N: (%f, %f, %f)", &o.N[0], &o.N[1], &o.N[2]) == 3, "N:");
    } else if (o.type == PTB_SCN_MESH) {                                 // TriangleMesh.h:143-167
        NEED(L.next(s), "is_centered: / has_csv:");
        if (s[0] == 'i' && s[1] == 's') { NEED(sscanf(s.c_str(), "is_centered: %u", &u) == 1, "is_centered:"); o.is_centered = u == 1; NEED(L.next(s), "has_csv:"); }
        NEED(sscanf(s.c_str(), "has_csv: %u", &u) == 1, "has_csv:"); o.has_csv = (int)u;
        NEED(L.next(s) && starts(s, "csv_file:"), "csv_file:");
        if (o.has_csv) copy_str(o.csv_file, after(s, "csv_file:"));
        o.display_edges = 0; o.interp_normals = 1;                       // TriMesh::init resets both (TriangleMesh.cpp:719-720)
    } else {                                                             // PointSet (PointSet.h:197-215): radius + file lines
        while (L.next(s)) if (starts(s, "NEW ") || starts(s, "fog_density:")) { L.unread(); break; }
    }
    return PTB_OK;
}

// Raytracer::load_scene (Raytracer.cpp:1149-1236)
int parse_scn(const char* path, const char* replaced, ptb_scn& sc) {
    FILE* f = fopen(path, "r");
    if (!f) return fail(PTB_ERR_INVALID, std::string("cannot open scene file '") + path + "'");
    Lines L;
    char buf[4096];
    while (fgets(buf, sizeof buf, f)) L.l.push_back(buf);
    fclose(f);
    sc.dir = dir_with_slash(path);
    ptb_scn_header& h = sc.h;
    memset(&h, 0, sizeof(h));
    h.nbframes = 1; h.nb_bounces = 3; h.envmap_intensity = 1;
    std::string s;
    unsigned a = 0, b = 0;
    NEED(L.next(s) && sscanf(s.c_str(), "W,H: %u, %u", &a, &b) == 2, "W,H:"); h.W = (int)a; h.H = (int)b;
    NEED(L.next(s) && sscanf(s.c_str(), "nrays: %u", &a) == 1, "nrays:"); h.nrays = (int)a;
    NEED(L.next(s), "nbframes: / Cam:");
    if (s[0] == 'n') { NEED(sscanf(s.c_str(), "nbframes: %u", &a) == 1, "nbframes:"); h.nbframes = (int)a; NEED(L.next(s), "Cam:"); }
    float* p = h.cam.position; float* d = h.cam.direction; float* up = h.cam.up;
    NEED(sscanf(s.c_str(), "Cam: (%f, %f, %f), (%f, %f, %f), (%f, %f, %f)", &p[0], &p[1], &p[2], &d[0], &d[1], &d[2], &up[0], &up[1], &up[2]) == 9, "Cam:");
    NEED(L.next(s) && sscanf(s.c_str(), "fov: %f", &h.cam.fov) == 1, "fov:");
    NEED(L.next(s) && sscanf(s.c_str(), "focus: %f", &h.cam.focus_distance) == 1, "focus:");
    NEED(L.next(s) && sscanf(s.c_str(), "aperture: %f", &h.cam.aperture) == 1, "aperture:");
    NEED(L.next(s) && sscanf(s.c_str(), "sigma_filter: %f", &h.sigma_filter) == 1, "sigma_filter:");
    NEED(L.next(s) && sscanf(s.c_str(), "gamma: %f", &h.gamma) == 1, "gamma:");
    NEED(L.next(s), "is_lenticular: / bounces:");
    if (sscanf(s.c_str(), "is_lenticular: %u", &a) == 1) {
        h.is_lenticular = (int)a;
        for (int k = 0; k < 8; k++) NEED(L.next(s), "lenticular / camera-array field");   // nb_images, max_angle, pixel_width, isArray, nbviewX/Y, maxSpacingX/Y
        NEED(L.next(s), "bounces:");
    }
    NEED(sscanf(s.c_str(), "bounces: %u", &a) == 1, "bounces:"); h.nb_bounces = (int)a;
    NEED(L.next(s), "has_denoiser: / intensite_lum:");
    if (sscanf(s.c_str(), "has_denoiser: %u", &a) == 1) { h.has_denoiser = (int)a; NEED(L.next(s), "intensite_lum:"); }
    NEED(sscanf(s.c_str(), "intensite_lum: %f", &h.intensite_lumiere) == 1, "intensite_lum:");
    NEED(L.next(s) && sscanf(s.c_str(), "intensite_envmap: %f", &h.envmap_intensity) == 1, "intensite_envmap:");
    NEED(L.next(s), "background: / nbobjects:");
    if (s[0] != 'n') { NEED(starts(s, "background:"), "background:"); copy_str(h.background, after(s, "background:")); NEED(L.next(s), "nbobjects:"); }
    NEED(sscanf(s.c_str(), "nbobjects: %u", &a) == 1, "nbobjects:"); h.n_objects = (int)a;
    sc.objects.resize(a);
    for (unsigned i = 0; i < a; i++) { int rc = parse_object(L, sc.objects[i], replaced); if (rc) return rc; }
    // trailing fog block: older files stop early; missing fields keep their zero defaults
    if (L.next(s) && sscanf(s.c_str(), "fog_density: %f", &h.fog_density) == 1 && L.next(s)) {
        if (sscanf(s.c_str(), "fog_absorption: %f", &h.fog_absorption) == 1) {
            if (L.next(s)) sscanf(s.c_str(), "fog_density_decay: %f", &h.fog_density_decay);
            if (L.next(s)) sscanf(s.c_str(), "fog_absorption_decay: %f", &h.fog_absorption_decay);
            if (L.next(s)) { if (sscanf(s.c_str(), "fog_type: %u", &a) == 1) h.fog_type = (int)a; }
        } else if (sscanf(s.c_str(), "fog_type: %u", &a) == 1) h.fog_type = (int)a;
        if (L.next(s) && sscanf(s.c_str(), "fog_phase_type: %u", &a) == 1) {
            h.fog_phase_type = (int)a;
            if (L.next(s)) sscanf(s.c_str(), "double_frustum_start_t: %f", &h.double_frustum_start_t);
        }
    }
    return PTB_OK;
}

std::string find_file(const ptb_scn& sc, const std::string& name) {     // as written (cwd-relative like the reference), else next to the .scn
    if (file_exists(name)) return name;
    if (!name.empty() && name[0] != '/' && file_exists(sc.dir + name)) return sc.dir + name;
    return name;
}

void fill_tex(ptb_tex& t, const Slot& sl, std::vector<float>& store, bool normals) {
    t.texels = nullptr; t.W = t.H = 0;
    memcpy(t.mult, sl.mult, sizeof t.mult);
    if (sl.file.empty()) return;
    int W = 0, H = 0;
    // a texture that fails to load stays a constant slot (Texture::loadColors leaves W = 0, BRDF.h:393-404)
    if (load_texture_values(sl.file.c_str(), normals ? 1 : 0, store, W, H) == PTB_OK) { t.texels = store.data(); t.W = W; t.H = H; }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C-ABI
// A C++ exception (std::bad_alloc on a header that asks for gigabytes, a length_error) must not unwind through the C boundary
#define PTB_IO_CATCH                                                                                      \
    catch (const std::bad_alloc&) { return fail(PTB_ERR_NOMEM, "out of host memory"); }                   \
    catch (const std::exception& e_) { return fail(PTB_ERR_INVALID, std::string("reader failed: ") + e_.what()); }

extern "C" {

const char* ptb_sceneio_last_error(void) { return g_err.c_str(); }

int ptb_image_load(const char* path, uint8_t** rgb, int32_t* W, int32_t* H) {
    try {
    if (!path || !rgb || !W || !H) return fail(PTB_ERR_INVALID, "image_load: null argument");
    Image im;
    int rc = load_image_flipped(path, im);
    if (rc) return rc;
    *rgb = (uint8_t*)malloc(im.rgb.size());
    if (!*rgb) return fail(PTB_ERR_NOMEM, "image_load: out of memory");
    memcpy(*rgb, im.rgb.data(), im.rgb.size());
    *W = im.W; *H = im.H;
    return PTB_OK;
    } PTB_IO_CATCH
}
void ptb_image_free(void* p) { free(p); }

int ptb_texture_load(const char* path, int kind, float** values, int32_t* W, int32_t* H) {
    try {
    if (!path || !values || !W || !H || (kind != 0 && kind != 1)) return fail(PTB_ERR_INVALID, "texture_load: bad argument");
    std::vector<float> v;
    int w = 0, h = 0;
    int rc = load_texture_values(path, kind, v, w, h);
    if (rc) return rc;
    *values = (float*)malloc(v.size() * sizeof(float));
    if (!*values) return fail(PTB_ERR_NOMEM, "texture_load: out of memory");
    memcpy(*values, v.data(), v.size() * sizeof(float));
    *W = w; *H = h;
    return PTB_OK;
    } PTB_IO_CATCH
}

// Yarns::Yarns(filename) (TriangleMesh.h:268-288): "%u\n" counts and "%f %f %f\n" points are whitespace-separated tokens to fscanf
int ptb_yarnfile_read(const char* path, float** A, float** B, float** R, int32_t* n_segments) {
    try {
    if (!path || !A || !B || !R || !n_segments) return fail(PTB_ERR_INVALID, "yarnfile_read: null argument");
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) return fail(PTB_ERR_INVALID, std::string("yarnfile_read: cannot read ") + path);
    buf.push_back(0);
    const char* p = reinterpret_cast<const char*>(buf.data());
    auto next_uint = [&](long& v) { char* e; v = strtol(p, &e, 10); if (e == p || v < 0) return false; p = e; return true; };
    auto next_float = [&](float& v) { char* e; v = strtof(p, &e); if (e == p) return false; p = e; return true; };
    long nbyarns;
    if (!next_uint(nbyarns)) return fail(PTB_ERR_INVALID, "yarnfile_read: no yarn count");
    std::vector<float> a, b;
    for (long y = 0; y < nbyarns; y++) {
        long nbpoints;
        if (!next_uint(nbpoints)) return fail(PTB_ERR_INVALID, "yarnfile_read: a yarn without its point count");
        float prev[3] = {0, 0, 0};
        for (long j = 0; j < nbpoints; j++) {
            float q[3];
            if (!next_float(q[0]) || !next_float(q[1]) || !next_float(q[2])) return fail(PTB_ERR_INVALID, "yarnfile_read: fewer points than announced");
            for (int k = 0; k < 3; k++) q[k] *= 50.f;                                   // Vector(x, y, z) * 50.f
            if (j > 0) { a.insert(a.end(), prev, prev + 3); b.insert(b.end(), q, q + 3); }
            memcpy(prev, q, sizeof q);
        }
        if (a.size() / 3 > (size_t)(1 << 25)) return fail(PTB_ERR_UNSUPPORTED, "yarnfile_read: more than 2^25 segments");
    }
    const size_t n = a.size() / 3;
    if (n == 0) return fail(PTB_ERR_INVALID, "yarnfile_read: no segment in the file");
    *A = (float*)malloc(3 * n * sizeof(float)); *B = (float*)malloc(3 * n * sizeof(float)); *R = (float*)malloc(n * sizeof(float));
    if (!*A || !*B || !*R) { free(*A); free(*B); free(*R); return fail(PTB_ERR_NOMEM, "yarnfile_read: out of memory"); }
    memcpy(*A, a.data(), 3 * n * sizeof(float)); memcpy(*B, b.data(), 3 * n * sizeof(float));
    for (size_t i = 0; i < n; i++) (*R)[i] = 0.1f;
    *n_segments = (int32_t)n;
    return PTB_OK;
    } PTB_IO_CATCH
}
void ptb_yarnfile_free(void* p) { free(p); }

int ptb_meshfile_read(const char* path, int load_textures, ptb_meshfile** out) {
    try {
    if (!path || !out) return fail(PTB_ERR_INVALID, "meshfile_read: null argument");
    ptb_meshfile* m = new ptb_meshfile();
    int rc = read_meshfile(path, load_textures, *m);
    if (rc) { delete m; return rc; }
    *out = m;
    return PTB_OK;
    } PTB_IO_CATCH
}
void ptb_meshfile_free(ptb_meshfile* m) { delete m; }
int ptb_meshfile_get(const ptb_meshfile* m, ptb_meshfile_info* o) {
    if (!m || !o) return fail(PTB_ERR_INVALID, "meshfile_get: null argument");
    o->vertices = m->vertices.data(); o->n_vertices = (int32_t)(m->vertices.size() / 3);
    o->normals = m->normals.data(); o->n_normals = (int32_t)(m->normals.size() / 3);
    o->uvs = m->uvs.data(); o->n_uvs = (int32_t)(m->uvs.size() / 2);
    o->vertex_colors = m->vertex_colors.data(); o->n_vertex_colors = (int32_t)(m->vertex_colors.size() / 3);
    o->tri = m->tri.data(); o->n_tri = (int32_t)(m->tri.size() / 10);
    o->n_groups = (int32_t)m->group_names.size();
    o->has_materials = m->has_materials ? 1 : 0;
    return PTB_OK;
}
int ptb_meshfile_group_name(const ptb_meshfile* m, int group, char name[PTB_PATH_MAX]) {
    if (!m || !name) return fail(PTB_ERR_INVALID, "meshfile_group_name: null argument");
    for (const auto& kv : m->group_names) if (kv.second == group) { copy_str(name, kv.first); return PTB_OK; }
    return fail(PTB_ERR_INVALID, "meshfile_group_name: no such group");
}
int ptb_meshfile_group_slot(const ptb_meshfile* m, int group, int kind, ptb_slot* out) {
    if (!m || !out || kind < 0 || kind >= PTB_N_KINDS) return fail(PTB_ERR_INVALID, "meshfile_group_slot: bad argument");
    if (!m->has_materials || group < 0 || group >= (int)m->slots[kind].size()) return fail(PTB_ERR_INVALID, "meshfile_group_slot: no such slot");
    const Slot& s = m->slots[kind][group];
    copy_str(out->file, s.file);
    memcpy(out->mult, s.mult, sizeof out->mult);
    return PTB_OK;
}

int ptb_scn_load(const char* path, const char* replaced_names, ptb_scn** out) {
    try {
    if (!path || !out) return fail(PTB_ERR_INVALID, "scn_load: null argument");
    ptb_scn* s = new ptb_scn();
    int rc = parse_scn(path, replaced_names, *s);
    if (rc) { delete s; return rc; }
    *out = s;
    return PTB_OK;
    } PTB_IO_CATCH
}
void ptb_scn_free(ptb_scn* s) { delete s; }
int ptb_scn_get_header(const ptb_scn* s, ptb_scn_header* out) {
    if (!s || !out) return fail(PTB_ERR_INVALID, "scn_get_header: null argument");
    *out = s->h;
    return PTB_OK;
}
int ptb_scn_get_object(const ptb_scn* s, int obj, ptb_scn_object* out) {
    if (!s || !out || obj < 0 || obj >= (int)s->objects.size()) return fail(PTB_ERR_INVALID, "scn_get_object: bad argument");
    *out = s->objects[obj].o;
    return PTB_OK;
}
int ptb_scn_get_slot(const ptb_scn* s, int obj, int kind, int idx, ptb_slot* out) {
    if (!s || !out || obj < 0 || obj >= (int)s->objects.size() || kind < 0 || kind >= PTB_N_KINDS) return fail(PTB_ERR_INVALID, "scn_get_slot: bad argument");
    const std::vector<Slot>& v = s->objects[obj].slots[kind];
    if (idx < 0 || idx >= (int)v.size()) return fail(PTB_ERR_INVALID, "scn_get_slot: no such slot");
    copy_str(out->file, v[idx].file.empty() ? std::string() : find_file(*s, v[idx].file));
    memcpy(out->mult, v[idx].mult, sizeof out->mult);
    return PTB_OK;
}

int ptb_scn_get_keyframes(const ptb_scn* s, int obj, int kind, float* frames, float* values, int cap) {
    if (!s || obj < 0 || obj >= (int)s->objects.size() || kind < 0 || kind > 2) return fail(PTB_ERR_INVALID, "scn_get_keyframes: bad argument");
    ptb::KeyTrack tracks[3];
    keys_to_tracks(s->objects[obj].keys, tracks);
    const ptb::KeyTrack& t = tracks[kind];
    const int n = (int)t.frames.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (frames) frames[i] = t.frames[i];
        if (values) std::copy(t.values.begin() + (size_t)i * t.width, t.values.begin() + (size_t)(i + 1) * t.width, values + (size_t)i * t.width);
    }
    return n;
}

int ptb_scn_save(const ptb_scn* s, const char* path) {
    try {
    if (!s || !path) return fail(PTB_ERR_INVALID, "scn_save: null argument");
    FILE* f = fopen(path, "w");
    if (!f) return fail(PTB_ERR_INVALID, std::string("cannot write '") + path + "'");
    const ptb_scn_header& h = s->h;
    fprintf(f, "W,H: %u, %u\nnrays: %u\nnbframes: %u\n", h.W, h.H, h.nrays, h.nbframes);
    fprintf(f, "Cam: (%f, %f, %f), (%f, %f, %f), (%f, %f, %f)\n", h.cam.position[0], h.cam.position[1], h.cam.position[2], h.cam.direction[0],
            h.cam.direction[1], h.cam.direction[2], h.cam.up[0], h.cam.up[1], h.cam.up[2]);
    fprintf(f, "fov: %f\nfocus: %f\naperture: %f\nsigma_filter: %f\ngamma: %f\n", h.cam.fov, h.cam.focus_distance, h.cam.aperture, h.sigma_filter, h.gamma);
    fprintf(f, "bounces: %u\nhas_denoiser: %u\nintensite_lum: %f\nintensite_envmap: %f\n", h.nb_bounces, h.has_denoiser, h.intensite_lumiere, h.envmap_intensity);
    if (h.background[0]) fprintf(f, "background: %s\n", h.background);
    fprintf(f, "nbobjects: %u\n", (unsigned)s->objects.size());
    static const char* kind_key[PTB_N_KINDS] = {"nb_textures", "nb_normalmaps", "nb_subsurfaces", "nb_specularmaps", "nb_alphamaps", "nb_expmaps", "nb_transpmaps", "nb_refrindexmaps"};
    for (const ScnObject& so : s->objects) {
        const ptb_scn_object& o = so.o;
        const char* head[] = {"NEW MESH", "NEW SPHERE", "NEW PLANE", "NEW POINTSET"};
        fprintf(f, "%s\nname: %s\nmiroir: %u\nghost: %u\n", head[o.type], o.name, o.miroir, o.ghost);
        const float* t = o.xform.translation; const float* r = o.xform.rotation; const float* c = o.xform.rotation_center;
        fprintf(f, "translation: (%f, %f, %f)\n", t[0], t[1], t[2]);
        fprintf(f, "rotation: (%f, %f, %f, %f, %f, %f, %f, %f, %f)\n", r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]);
        fprintf(f, "center: (%f, %f, %f)\nscale: %f\ndisplay_edges: %u\ninterp_normals: %u\nflip_normals: %u\n", c[0], c[1], c[2], o.xform.scale,
                o.display_edges, o.interp_normals, o.flip_normals);
        {   // Geometry.h:466-476: nb_transforms = translation_keyframes.size(), then the scale, translation and rotation rows
            ptb::KeyTrack tk[3];
            keys_to_tracks(so.keys, tk);
            fprintf(f, "nb_transforms: %u\n", (unsigned)tk[1].frames.size());
            for (size_t i = 0; i < tk[0].frames.size(); i++) fprintf(f, "%f %f\n", tk[0].frames[i], tk[0].values[i]);
            for (size_t i = 0; i < tk[1].frames.size(); i++) fprintf(f, "%f %f, %f, %f\n", tk[1].frames[i], tk[1].values[3 * i], tk[1].values[3 * i + 1], tk[1].values[3 * i + 2]);
            for (size_t i = 0; i < tk[2].frames.size(); i++) {
                const float* m = &tk[2].values[9 * i];
                fprintf(f, "%f %f, %f, %f, %f, %f, %f, %f, %f, %f\n", tk[2].frames[i], m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
            }
        }
        for (int k = 0; k < PTB_N_KINDS; k++) {
            fprintf(f, "%s: %u\n", kind_key[k], (unsigned)so.slots[k].size());
            for (const Slot& sl : so.slots[k]) {
                fprintf(f, "texture: %s\n", sl.file.empty() ? "Null" : sl.file.c_str());
                if (k == PTB_KIND_TRANSP || k == PTB_KIND_REFR) fprintf(f, "multiplier: %f)\n", sl.mult[0]);
                else fprintf(f, "multiplier: (%f, %f, %f)\n", sl.mult[0], sl.mult[1], sl.mult[2]);
            }
        }
        if (o.type == PTB_SCN_SPHERE) fprintf(f, "is_envmap: %u\nenvmapfilename: %s\nO: (%f, %f, %f)\nR: %f\n", o.is_envmap, o.envmap, o.O[0], o.O[1], o.O[2], o.R);
        else if (o.type == PTB_SCN_PLANE) fprintf(f, "Point: (%f, %f, %f)\nN: (%f, %f, %f)\n", o.A[0], o.A[1], o.A[2], o.N[0], o.N[1], o.N[2]);
        else if (o.type == PTB_SCN_MESH) fprintf(f, "is_centered: %u\nhas_csv: %u\ncsv_file: %s\n", o.is_centered, o.has_csv, o.csv_file);
    }
    fprintf(f, "fog_density: %f\nfog_absorption: %f\nfog_density_decay: %f\nfog_absorption_decay: %f\nfog_type: %u\nfog_phase_type: %u\ndouble_frustum_start_t: %f\n",
            h.fog_density, h.fog_absorption, h.fog_density_decay, h.fog_absorption_decay, h.fog_type, h.fog_phase_type, h.double_frustum_start_t);
    fclose(f);
    return PTB_OK;
    } PTB_IO_CATCH
}

int ptb_load_scene(ptb_ctx* ctx, const char* path, const char* replaced_names, ptb_camera* cam, ptb_params* params) {
    try {
    if (!ctx || !path) return fail(PTB_ERR_INVALID, "load_scene: null argument");
    ptb_scn sc;
    int rc = parse_scn(path, replaced_names, sc);
    if (rc) return rc;
    const ptb_scn_header& h = sc.h;
    if (h.is_lenticular) return fail(PTB_ERR_UNSUPPORTED, "load_scene: lenticular cameras are not rendered");
    if (sc.objects.size() < 2 || sc.objects[0].o.type != PTB_SCN_SPHERE) return fail(PTB_ERR_INVALID, "load_scene: object 0 must be the light sphere");
    for (size_t i = 0; i < sc.objects.size(); i++) {
        const ScnObject& so = sc.objects[i];
        const ptb_scn_object& o = so.o;
        int flags = (o.miroir ? PTB_OBJ_MIRROR : 0) | (o.flip_normals ? PTB_OBJ_FLIP_NORMALS : 0) | (o.interp_normals ? 0 : PTB_OBJ_FLAT_NORMALS) |
                    (o.ghost ? PTB_OBJ_GHOST : 0);
        int id = -1;
        if (o.type == PTB_SCN_SPHERE) {
            rc = ptb_add_sphere(ctx, o.O, o.R, &o.xform, flags, &id);
            if (!rc && o.is_envmap) {
                if (i != 1) return fail(PTB_ERR_UNSUPPORTED, "load_scene: environment maps are only rendered on object 1 (the dome)");
                Image im;
                if ((rc = load_image_flipped(find_file(sc, o.envmap).c_str(), im))) return rc;
                rc = ptb_set_envmap(ctx, im.rgb.data(), im.W, im.H);
            }
        } else if (o.type == PTB_SCN_PLANE) {
            rc = ptb_add_plane(ctx, o.A, o.N, &o.xform, flags, &id);
        } else if (o.type == PTB_SCN_MESH) {
            ptb_meshfile mf;
            if ((rc = read_meshfile(find_file(sc, o.name).c_str(), 0, mf))) return rc;
            if (!mf.vertex_colors.empty()) return fail(PTB_ERR_UNSUPPORTED, "load_scene: per-vertex colours are not rendered");
            ptb_mesh m;
            memset(&m, 0, sizeof m);
            m.vertices = mf.vertices.data(); m.n_vertices = (int32_t)(mf.vertices.size() / 3);
            m.normals = mf.normals.data(); m.n_normals = (int32_t)(mf.normals.size() / 3);
            m.uvs = mf.uvs.data(); m.n_uvs = (int32_t)(mf.uvs.size() / 2);
            m.tri = mf.tri.data(); m.n_tri = (int32_t)(mf.tri.size() / 10);
            m.scaling = 1.f; m.center = o.is_centered;            // TriangleMesh.h:165: init(scene, name, 1., (0,0,0), ..., is_centered, rotation_center)
            rc = ptb_add_mesh(ctx, &m, &o.xform, flags, &id);
        } else return fail(PTB_ERR_UNSUPPORTED, "load_scene: PointSet objects are not rendered");
        if (rc) return fail(rc, std::string("load_scene: object ") + std::to_string(i) + ": " + ptb_last_error(ctx));
        {   // the keyframe maps travel with the object; ptb_set_frame + ptb_commit place it (Scene::prepare_render -> build_matrix(frame))
            ptb::KeyTrack tk[3];
            keys_to_tracks(so.keys, tk);
            for (int k = 0; k < 3; k++)
                if (!tk[k].empty() && (rc = ptb_set_keyframes(ctx, id, k, tk[k].frames.data(), tk[k].values.data(), (int)tk[k].frames.size()))) return fail(rc, "load_scene: keyframes");
        }
        // per-group slots: a kind is present for group g iff g < its vector's length (Object::queryMaterial, Geometry.h:399-445)
        size_t ng = 0;
        for (int k = 0; k < PTB_N_KINDS; k++) ng = std::max(ng, so.slots[k].size());
        for (size_t g = 0; g < ng; g++) {
            ptb_material mat;
            memset(&mat, 0, sizeof mat);
            std::vector<float> store[PTB_N_KINDS];
            struct { int kind; uint32_t bit; ptb_tex* t; } map[] = {
                {PTB_KIND_KD, PTB_SLOT_KD, &mat.Kd}, {PTB_KIND_KS, PTB_SLOT_KS, &mat.Ks}, {PTB_KIND_NE, PTB_SLOT_NE, &mat.Ne}, {PTB_KIND_TRANSP, PTB_SLOT_TRANSP, &mat.transp},
                {PTB_KIND_REFR, PTB_SLOT_REFR, &mat.refr}, {PTB_KIND_NORMAL, PTB_SLOT_NORMAL, &mat.normal}, {PTB_KIND_ALPHA, PTB_SLOT_ALPHA, &mat.alpha},
                // Object::subsurface: rendered on triangle meshes (Raytracer.cpp:318-406); ptb_set_group_material refuses a non-zero one elsewhere
                {PTB_KIND_SUBSURF, PTB_SLOT_KSUB, &mat.Ksub}};
            for (auto& e : map) {
                if (g >= so.slots[e.kind].size()) continue;
                Slot sl = so.slots[e.kind][g];
                if (!sl.file.empty()) sl.file = find_file(sc, sl.file);
                mat.present |= e.bit;
                fill_tex(*e.t, sl, store[e.kind], e.kind == PTB_KIND_NORMAL);
            }
            if ((rc = ptb_set_group_material(ctx, id, (int)g, &mat))) return fail(rc, std::string("load_scene: material: ") + ptb_last_error(ctx));
        }
    }
    if ((rc = ptb_set_light(ctx, h.intensite_lumiere, h.envmap_intensity))) return rc;
    {   // Scene::fog_* as Raytracer::load_scene reads them (Raytracer.cpp:1216-1231); phase_aniso is not stored in .scn files
        ptb_fog fog = {h.fog_density, h.fog_absorption, h.fog_density_decay, h.fog_absorption_decay, h.fog_type, h.fog_phase_type, 0.f};
        if ((rc = ptb_set_fog(ctx, &fog))) return fail(rc, std::string("load_scene: ") + ptb_last_error(ctx));
    }
    if (h.background[0]) {   // Scene::load_background(bg, gamma) (Raytracer.cpp:1208-1209, Geometry.h:1355-1363)
        Image im;
        if ((rc = load_image_flipped(find_file(sc, h.background).c_str(), im))) return rc;
        std::vector<float> bg(im.rgb.size());
        for (size_t i = 0; i < bg.size(); i++) bg[i] = (float)(std::pow(im.rgb[i] / 255., (double)h.gamma) * 196964.699);
        if ((rc = ptb_set_background(ctx, bg.data(), im.W, im.H))) return rc;
    } else ptb_set_background(ctx, nullptr, 0, 0);
    if (cam) *cam = h.cam;
    if (params) {
        memset(params, 0, sizeof *params);
        params->W = h.W; params->H = h.H; params->nrays = h.nrays; params->nb_bounces = h.nb_bounces;
        params->sigma_filter = h.sigma_filter; params->gamma = h.gamma; params->shard_count = 1;
    }
    return PTB_OK;
    } PTB_IO_CATCH
}

}  // extern "C"
