// tgaimage.cpp — TGA read (types 2/3 raw, 10/11 RLE; 8/24/32 bpp; origin flags) and write (RLE or raw,
// bottom-left origin by default).  Behaviour follows reference src/tgaimage.cpp:43-246 (which is itself
// ssloy/tinyrenderer's codec): the 18-byte header is not followed by an id-field skip, bottom-origin files are
// flipped so that row 0 is the top row of the file's image, and the writer emits the TRUEVISION-XFILE footer.
#include "tgaimage.h"

#include <cstdio>
#include <cstring>

namespace
{
#pragma pack(push, 1)
struct Header
{
    std::uint8_t  idlength, colormaptype, datatypecode;
    std::uint16_t colormaporigin, colormaplength;
    std::uint8_t  colormapdepth;
    std::uint16_t x_origin, y_origin, width, height;
    std::uint8_t  bitsperpixel, imagedescriptor;
};
#pragma pack(pop)

struct File
{
    FILE* f;
    explicit File(FILE* ff) : f(ff) {}
    ~File() { if (f) fclose(f); }
};
}  // namespace

bool TGAImage::ReadTgaFile(const std::string& filename)
{
    File in(fopen(filename.c_str(), "rb"));
    if (!in.f) return false;
    Header h;
    if (fread(&h, sizeof h, 1, in.f) != 1) return false;
    m_Width = h.width;
    m_Height = h.height;
    m_Bytespp = h.bitsperpixel >> 3;
    if (m_Width <= 0 || m_Height <= 0 || (m_Bytespp != GRAYSCALE && m_Bytespp != RGB && m_Bytespp != RGBA))
        return false;
    size_t nbytes = (size_t)m_Bytespp * m_Width * m_Height;
    m_Data.assign(nbytes, 0);
    if (h.datatypecode == 2 || h.datatypecode == 3)
    {
        if (fread(m_Data.data(), 1, nbytes, in.f) != nbytes) return false;
    }
    else if (h.datatypecode == 10 || h.datatypecode == 11)
    {
        size_t       pixels = (size_t)m_Width * m_Height, cur = 0, byte = 0;
        std::uint8_t px[4];
        while (cur < pixels)
        {
            int c = fgetc(in.f);
            if (c == EOF) return false;
            int  count = (c & 127) + 1;
            bool run = c >= 128;
            if (run && fread(px, 1, m_Bytespp, in.f) != (size_t)m_Bytespp) return false;
            for (int i = 0; i < count; ++i)
            {
                if (!run && fread(px, 1, m_Bytespp, in.f) != (size_t)m_Bytespp) return false;
                if (++cur > pixels) return false;
                for (int t = 0; t < m_Bytespp; ++t) m_Data[byte++] = px[t];
            }
        }
    }
    else
        return false;
    if (!(h.imagedescriptor & 0x20)) FlipVertically();
    if (h.imagedescriptor & 0x10) FlipHorizontally();
    return true;
}

bool TGAImage::WriteTgaFile(const std::string& filename, bool vFlip, bool rle) const
{
    static const std::uint8_t tail[26] = { 0, 0, 0, 0, 0, 0, 0, 0, 'T', 'R', 'U', 'E', 'V', 'I', 'S', 'I', 'O',
                                           'N', '-', 'X', 'F', 'I', 'L', 'E', '.', '\0' };
    File out(fopen(filename.c_str(), "wb"));
    if (!out.f) return false;
    Header h;
    memset(&h, 0, sizeof h);
    h.bitsperpixel = (std::uint8_t)(m_Bytespp << 3);
    h.width = (std::uint16_t)m_Width;
    h.height = (std::uint16_t)m_Height;
    h.datatypecode = (m_Bytespp == GRAYSCALE) ? (rle ? 11 : 3) : (rle ? 10 : 2);
    h.imagedescriptor = vFlip ? 0x00 : 0x20;
    if (fwrite(&h, sizeof h, 1, out.f) != 1) return false;
    if (!rle)
    {
        if (fwrite(m_Data.data(), 1, m_Data.size(), out.f) != m_Data.size()) return false;
    }
    else
    {
        // Greedy packets of at most 128 pixels: a raw packet ends before the first repeated pair, a run packet
        // ends at the first differing pixel — the same packetisation as the reference's encoder
        // (tgaimage.cpp:249-300), so files are byte-comparable with the oracle's.
        const size_t npix = (size_t)m_Width * m_Height, bpp = m_Bytespp;
        size_t       cur = 0;
        while (cur < npix)
        {
            auto   same = [&](size_t a, size_t b) { return memcmp(&m_Data[a * bpp], &m_Data[b * bpp], bpp) == 0; };
            size_t len = 1;
            bool   raw = true;
            while (cur + len < npix && len < 128)
            {
                bool eq = same(cur + len - 1, cur + len);
                if (len == 1) raw = !eq;
                if (raw && eq) { --len; break; }
                if (!raw && !eq) break;
                ++len;
            }
            fputc(raw ? (int)len - 1 : (int)len + 127, out.f);
            if (fwrite(&m_Data[cur * bpp], 1, raw ? len * bpp : bpp, out.f) != (raw ? len * bpp : bpp)) return false;
            cur += len;
        }
    }
    return fwrite(tail, 1, sizeof tail, out.f) == sizeof tail;
}

TGAColor TGAImage::Get(int x, int y) const
{
    if (m_Data.empty() || x < 0 || y < 0 || x >= m_Width || y >= m_Height) return TGAColor(0, 0, 0);
    return TGAColor(m_Data.data() + ((size_t)x + (size_t)y * m_Width) * m_Bytespp, (std::uint8_t)m_Bytespp);
}

void TGAImage::Set(int x, int y, const TGAColor& c)
{
    if (m_Data.empty() || x < 0 || y < 0 || x >= m_Width || y >= m_Height) return;
    memcpy(m_Data.data() + ((size_t)x + (size_t)y * m_Width) * m_Bytespp, c.bgra, m_Bytespp);
}

void TGAImage::FlipHorizontally()
{
    for (int j = 0; j < m_Height; ++j)
        for (int i = 0; i < m_Width / 2; ++i)
        {
            TGAColor a = Get(i, j), b = Get(m_Width - 1 - i, j);
            Set(i, j, b);
            Set(m_Width - 1 - i, j, a);
        }
}

void TGAImage::FlipVertically()
{
    size_t                    line = (size_t)m_Width * m_Bytespp;
    std::vector<std::uint8_t> tmp(line);
    for (int j = 0; j < m_Height / 2; ++j)
    {
        std::uint8_t* a = m_Data.data() + (size_t)j * line;
        std::uint8_t* b = m_Data.data() + (size_t)(m_Height - 1 - j) * line;
        memcpy(tmp.data(), a, line);
        memcpy(a, b, line);
        memcpy(b, tmp.data(), line);
    }
}
