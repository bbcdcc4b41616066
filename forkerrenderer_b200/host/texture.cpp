// texture.cpp — lazy upload of a Texture to the device, and host-side sampling (reference src/texture.h:41-145).
#include "texture.h"

#include "forkergl.h"
#include "forkergl_b200.h"

int Texture::DeviceId() const
{
    if (m_DeviceId < 0)
    {
        ForkerGL::Check(fgl_upload_texture(ForkerGL::Context(), m_Image.Buffer(), m_Image.GetWidth(),
                                           m_Image.GetHeight(), m_Image.GetBytespp(), (int)m_WrapMode,
                                           (int)m_FilterMode, &m_DeviceId),
                        "upload texture");
    }
    return m_DeviceId;
}

// ---- host-side sampling ------------------------------------------------------------------------------------------------
// Wrap (texture.h:62-83): NoWrap leaves the coordinate alone (texels outside the image read as black, tgaimage.cpp:304-311),
// Repeat keeps the fraction, MirroredRepeat flips every other period, ClampToEdge clamps to [0, 1].  The image is
// addressed as [0, W - 0.001) x [0, H - 0.001); Nearest floors, Linear blends the four texels around the sample with
// texel centres at half-integers and clamps the taps unless the mode is NoWrap (texture.h:86-132).  A grey image keeps its
// value in the blue channel, which is what SampleFloat returns.
static int TruncX86(Float f) { return (f > -2147483904.f && f < 2147483648.f) ? (int)f : (int)0x80000000; }

Vector3f Texture::filtered(Float u, Float v) const
{
    if (m_WrapMode == Repeat) u = u - std::floor(u), v = v - std::floor(v);
    else if (m_WrapMode == MirroredRepeat)
    {
        int   xi = TruncX86(std::floor(u)), yi = TruncX86(std::floor(v));
        Float rx = u - (Float)xi, ry = v - (Float)yi;
        u = xi % 2 == 0 ? rx : 1.f - rx;
        v = yi % 2 == 0 ? ry : 1.f - ry;
    }
    else if (m_WrapMode == ClampToEdge) u = Clamp(u, 0.f, 1.f), v = Clamp(v, 0.f, 1.f);
    auto texel = [&](int x, int y) {
        TGAColor c = m_Image.Get(x, y);
        return Vector3f((Float)c.r(), (Float)c.g(), (Float)c.b());
    };
    const Float w = (Float)((double)m_Width - 0.001), h = (Float)((double)m_Height - 0.001);
    if (m_FilterMode == Nearest) return texel(TruncX86(std::floor(u * w)), TruncX86(std::floor(v * h)));
    const Float px = u * w, py = v * h;
    const Float lx = std::floor(px - 0.5f), ly = std::floor(py - 0.5f);
    const Float tx = px - (lx + 0.5f), ty = py - (ly + 0.5f);
    int x0 = TruncX86(lx), y0 = TruncX86(ly), x1 = TruncX86(lx + 1.f), y1 = TruncX86(ly + 1.f);
    if (m_WrapMode != NoWrap)
    {
        x0 = Clamp(x0, 0, m_Width - 1), x1 = Clamp(x1, 0, m_Width - 1);
        y0 = Clamp(y0, 0, m_Height - 1), y1 = Clamp(y1, 0, m_Height - 1);
    }
    auto lerp = [](Float t, const Vector3f& a, const Vector3f& b) { return a * (1 - t) + b * t; };  // geometry.h:912-916
    return lerp(ty, lerp(tx, texel(x0, y0), texel(x1, y0)), lerp(tx, texel(x0, y1), texel(x1, y1)));
}

Color3 Texture::Sample(const Vector2f& coord) const { return filtered(coord.x, coord.y) / 255.f; }
Float  Texture::SampleFloat(const Vector2f& coord) const { return filtered(coord.x, coord.y).z / 255.f; }
