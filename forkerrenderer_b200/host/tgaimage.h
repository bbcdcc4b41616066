// tgaimage.h — TGA codec of the drop-in facade (same public names as reference src/tgaimage.h:29-146,
// fresh implementation).  Texel memory order is B,G,R[,A] (or one grey byte), row 0 first; Get() outside the
// image returns black and Set() outside is ignored (reference tgaimage.cpp:304-317) — texture sampling relies
// on exactly that.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

struct TGAColor
{
    std::uint8_t bgra[4] = { 0, 0, 0, 0 };
    std::uint8_t bytespp = 0;

    TGAColor() = default;
    TGAColor(std::uint8_t R, std::uint8_t G, std::uint8_t B) : bgra{ B, G, R, 0 }, bytespp(3) {}
    TGAColor(std::uint8_t R, std::uint8_t G, std::uint8_t B, std::uint8_t A) : bgra{ B, G, R, A }, bytespp(4) {}
    TGAColor(std::uint8_t v) : bgra{ v, 0, 0, 0 }, bytespp(1) {}
    TGAColor(const std::uint8_t* p, std::uint8_t bpp) : bytespp(bpp)
    {
        for (int i = 0; i < bpp; ++i) bgra[i] = p[i];
    }
    std::uint8_t  b() const { return bgra[0]; }
    std::uint8_t  g() const { return bgra[1]; }
    std::uint8_t  r() const { return bgra[2]; }
    std::uint8_t  a() const { return bgra[3]; }
    std::uint8_t  operator[](int i) const { return bgra[i]; }
    std::uint8_t& operator[](int i) { return bgra[i]; }
};

class TGAImage
{
public:
    enum Format { GRAYSCALE = 1, RGB = 3, RGBA = 4 };

    TGAImage() = default;
    TGAImage(int w, int h, int bpp) : m_Data((size_t)w * h * bpp, 0), m_Width(w), m_Height(h), m_Bytespp(bpp) {}

    bool ReadTgaFile(const std::string& filename);
    bool WriteTgaFile(const std::string& filename, bool vFlip = true, bool rle = true) const;

    TGAColor Get(int x, int y) const;
    void     Set(int x, int y, const TGAColor& c);

    void FlipHorizontally();
    void FlipVertically();

    int                 GetWidth() const { return m_Width; }
    int                 GetHeight() const { return m_Height; }
    int                 GetBytespp() const { return m_Bytespp; }
    std::uint8_t*       Buffer() { return m_Data.data(); }
    const std::uint8_t* Buffer() const { return m_Data.data(); }
    void                Clear() { m_Data.assign(m_Data.size(), 0); }

private:
    std::vector<std::uint8_t> m_Data;
    int                       m_Width = 0;
    int                       m_Height = 0;
    int                       m_Bytespp = 0;
};
