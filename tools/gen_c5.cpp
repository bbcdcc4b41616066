// gen_c5.cpp — writes the synthetic high-triangle-count workload C5 of SURVEY.md §8(d) as files the REFERENCE can load:
// an OBJ + MTL + three TGA textures + a .scene file in the reference's grammar (reference src/model.cpp:143-255, 258-408,
// src/scene.cpp:54-171, src/tgaimage.cpp:43-103).  Nothing here is renderer code; it is the input generator of a benchmark.
//
//   gen_c5 <quads per side> <width> <height> <asset dir> <scene file> [ssao on|off] [mode deferred|forward]
//
// Mesh: a quads x quads height field over x, z in [-1, 1],
//     y(x, z) = 0.15 * sum_{k=1..4} 2^-k * sin(2^k pi x + phi_k) * cos(2^k pi z + psi_k),
// phi_k, psi_k = 2 pi * draw / 2^32 with the draws taken from std::mt19937(20261017) in the order phi_1, psi_1, phi_2, ...;
// analytic normals; vt = (x, z) * 8 (exercises the Repeat wrap mode); one quad face per cell ("f a/a/a b/b/b c/c/c d/d/d",
// which the loaders fan into two triangles); one material with 256 x 256 procedural map_Kd (24 bit), map_Pr and map_Pm
// (8 bit grey), which makes the model PBR (model.cpp:58-66).  quads = 2237 gives 10 008 338 triangles.
// Numbers are printed with the shortest representation that round-trips to the same float (std::to_chars), so the text
// is a deterministic function of the arguments and every loader reads back identical floats.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace
{
struct Out
{
    FILE*             f;
    std::vector<char> buf;
    size_t            n = 0;
    explicit Out(const std::string& path) : f(fopen(path.c_str(), "wb")), buf(1 << 22) {}
    ~Out()
    {
        flush();
        if (f) fclose(f);
    }
    void flush()
    {
        if (f && n) fwrite(buf.data(), 1, n, f);
        n = 0;
    }
    char* room(size_t k)
    {
        if (n + k > buf.size()) flush();
        return buf.data() + n;
    }
    void str(const char* s)
    {
        size_t k = strlen(s);
        memcpy(room(k), s, k);
        n += k;
    }
    void num(float v)
    {
        char* p = room(32);
        auto  r = std::to_chars(p, p + 32, v);
        n += (size_t)(r.ptr - p);
    }
    void num(unsigned v)
    {
        char* p = room(16);
        auto  r = std::to_chars(p, p + 16, v);
        n += (size_t)(r.ptr - p);
    }
    void ch(char c)
    {
        *room(1) = c;
        ++n;
    }
};

// uncompressed TGA, bottom-left origin (imagedescriptor 0), 1 or 3 bytes per texel (grey / B,G,R)
bool write_tga(const std::string& path, int w, int h, int bpp, const std::vector<uint8_t>& texels)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    uint8_t hd[18] = { 0 };
    hd[2] = bpp == 1 ? 3 : 2;
    hd[12] = (uint8_t)(w & 255), hd[13] = (uint8_t)(w >> 8), hd[14] = (uint8_t)(h & 255), hd[15] = (uint8_t)(h >> 8);
    hd[16] = (uint8_t)(bpp * 8);
    fwrite(hd, 1, 18, f);
    fwrite(texels.data(), 1, texels.size(), f);
    fclose(f);
    return true;
}

uint32_t lcg(uint32_t& s)
{
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}
}  // namespace

int main(int argc, char** argv)
{
    if (argc < 6)
    {
        fprintf(stderr, "usage: gen_c5 <quads> <width> <height> <asset dir> <scene file> [ssao on|off] [mode deferred|forward]\n");
        return 2;
    }
    const int         Q = atoi(argv[1]), W = atoi(argv[2]), H = atoi(argv[3]);
    const std::string assets = argv[4], scene = argv[5];
    const std::string ssao = argc > 6 ? argv[6] : "on", mode = argc > 7 ? argv[7] : "deferred";
    if (Q < 1 || Q > 8192 || W < 1 || H < 1) return 2;
    const std::string name = "c5_" + std::to_string(Q), dir = assets + "/obj/" + name;
    if (system(("mkdir -p '" + dir + "'").c_str()) != 0) return 3;

    std::mt19937 rng(20261017u);
    double       phi[4], psi[4];
    const double kPi = 3.14159265358979323846;
    for (int k = 0; k < 4; ++k)
    {
        phi[k] = 2.0 * kPi * (double)rng() / 4294967296.0;
        psi[k] = 2.0 * kPi * (double)rng() / 4294967296.0;
    }
    const int N = Q + 1;
    {
        Out o(dir + "/field.obj");
        if (!o.f) return 3;
        o.str("# C5 synthetic height field (tools/gen_c5.cpp)\nmtllib field.mtl\ng field\nusemtl fieldmat\n");
        std::vector<float> nx((size_t)N * N), ny((size_t)N * N), nz((size_t)N * N);
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
            {
                const double x = -1.0 + 2.0 * (double)i / Q, z = -1.0 + 2.0 * (double)j / Q;
                double       y = 0, dx = 0, dz = 0;
                for (int k = 1; k <= 4; ++k)
                {
                    const double a = 0.15 * std::ldexp(1.0, -k), w = std::ldexp(1.0, k) * kPi;
                    const double sx = std::sin(w * x + phi[k - 1]), cx = std::cos(w * x + phi[k - 1]);
                    const double sz = std::sin(w * z + psi[k - 1]), cz = std::cos(w * z + psi[k - 1]);
                    y += a * sx * cz, dx += a * w * cx * cz, dz -= a * w * sx * sz;
                }
                const double len = std::sqrt(dx * dx + 1.0 + dz * dz);
                const size_t v = (size_t)j * N + i;
                nx[v] = (float)(-dx / len), ny[v] = (float)(1.0 / len), nz[v] = (float)(-dz / len);
                o.str("v "), o.num((float)x), o.ch(' '), o.num((float)y), o.ch(' '), o.num((float)z), o.ch('\n');
            }
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i)
            {
                const double x = -1.0 + 2.0 * (double)i / Q, z = -1.0 + 2.0 * (double)j / Q;
                o.str("vt "), o.num((float)(x * 8.0)), o.ch(' '), o.num((float)(z * 8.0)), o.ch('\n');
            }
        for (size_t v = 0; v < (size_t)N * N; ++v) o.str("vn "), o.num(nx[v]), o.ch(' '), o.num(ny[v]), o.ch(' '), o.num(nz[v]), o.ch('\n');
        for (int j = 0; j < Q; ++j)
            for (int i = 0; i < Q; ++i)
            {
                const unsigned a = (unsigned)(j * N + i) + 1, c[4] = { a, a + (unsigned)N, a + (unsigned)N + 1, a + 1 };
                o.ch('f');
                for (unsigned v : c) o.ch(' '), o.num(v), o.ch('/'), o.num(v), o.ch('/'), o.num(v);
                o.ch('\n');
            }
    }
    {
        FILE* f = fopen((dir + "/field.mtl").c_str(), "w");
        if (!f) return 3;
        fprintf(f, "newmtl fieldmat\nKa 0.3 0.3 0.3\nKd 0.8 0.6 0.4\nKs 0.5 0.5 0.5\nKe 0 0 0\nPr 0.6\nPm 0.2\n"
                   "map_Kd field_albedo.tga\nmap_Pr field_roughness.tga\nmap_Pm field_metalness.tga\n");
        fclose(f);
    }
    {
        const int            T = 256;
        std::vector<uint8_t> alb((size_t)T * T * 3), rough((size_t)T * T), metal((size_t)T * T);
        uint32_t             s = 20261017u;
        for (int y = 0; y < T; ++y)
            for (int x = 0; x < T; ++x)
            {
                const double u = (double)x / T, v = (double)y / T;
                const double base = 0.5 + 0.25 * std::sin(2 * kPi * 3 * u) * std::cos(2 * kPi * 2 * v);
                const size_t t = (size_t)y * T + x;
                for (int c = 0; c < 3; ++c)
                {
                    double val = base + 0.1 * c + 0.2 * ((double)(lcg(s) & 0xffff) / 65536.0) - 0.1;
                    alb[t * 3 + c] = (uint8_t)(255.0 * std::fmin(1.0, std::fmax(0.0, val)));
                }
                rough[t] = (uint8_t)(255.0 * (0.35 + 0.5 * (0.5 + 0.5 * std::sin(2 * kPi * 5 * u + 1.0) * std::sin(2 * kPi * 4 * v))));
                metal[t] = (uint8_t)(((x / 32 + y / 32) & 1) ? 200 : 30);
            }
        if (!write_tga(dir + "/field_albedo.tga", T, T, 3, alb) || !write_tga(dir + "/field_roughness.tga", T, T, 1, rough) ||
            !write_tga(dir + "/field_metalness.tga", T, T, 1, metal))
            return 3;
    }
    {
        FILE* f = fopen(scene.c_str(), "w");
        if (!f) return 3;
        fprintf(f,
                "# C5 (SURVEY.md 8d): %d x %d-quad height field = %lld triangles + the ground plane, written by tools/gen_c5.cpp\n"
                "mode %s\nscreen %d %d\nssaa off 2\nssao %s\nshadow on\nlight point 2 5 5 2 2 2\ncamera persp -1 1 1 0 0 -1\n"
                "model obj/plane/plane.obj false false 0 -1 -1 0 3\nmodel obj/%s/field.obj true false 0 -0.5 -1 0 2.5\n",
                Q, Q, 2LL * Q * Q, mode.c_str(), W, H, ssao.c_str(), name.c_str());
        fclose(f);
    }
    return 0;
}
