// buffer.cpp — Buffer1f/Buffer3f facade: host mirror synchronisation, 8-bit image generation
// (reference src/buffer.cpp:20-32,113-126) and the post-processing entry points (src/buffer.cpp:35-98,140-203),
// which run on the device through fgl_blur.
#include "buffer.h"

#include <stdexcept>

#include "forkergl.h"
#include "forkergl_b200.h"

void Buffer::pull() const
{
    if (m_HostValid) return;
    if (m_Plane < 0) throw std::runtime_error("Buffer: access to an empty (default-constructed) buffer");
    m_Host.resize((size_t)m_Width * m_Height * m_Channels);
    ForkerGL::Check(fgl_read_plane(ForkerGL::Context(), m_Plane, m_Host.data(), m_Host.size() * sizeof(Float)),
                    "Buffer read");
    m_HostValid = true;
}

void Buffer::push()
{
    if (!m_HostDirty || m_Plane < 0) return;
    ForkerGL::Check(fgl_write_plane(ForkerGL::Context(), m_Plane, m_Host.data(), m_Host.size() * sizeof(Float)),
                    "Buffer write");
    m_HostDirty = false;
}

// Standalone buffers (not one of ForkerGL's static planes) are not needed by any pass of Render::Render; they are
// provided as host-initialised mirrors of the AO / frame planes' element type for API completeness.
Buffer1f::Buffer1f(int w, int h, InitType type) : Buffer(w, h, -1, 1)
{
    Float v = type == Zero ? 0.f : type == One ? 1.f : type == MaxPositive ? 3.402823466e+38f : 1.175494351e-38f;
    m_Host.assign((size_t)w * h, v);
    m_HostValid = true;
}

Buffer3f::Buffer3f(int w, int h, InitType type) : Buffer(w, h, -1, 3)
{
    Float v = type == Zero ? 0.f : type == One ? 1.f : type == MaxPositive ? 3.402823466e+38f : 1.175494351e-38f;
    m_Host.assign((size_t)w * h * 3, v);
    m_HostValid = true;
}

TGAImage Buffer1f::GenerateImage(bool inverseColor) const
{
    pull();
    TGAImage image(m_Width, m_Height, TGAImage::GRAYSCALE);
    for (int x = 0; x < m_Width; ++x)
        for (int y = 0; y < m_Height; ++y)
        {
            Float val = inverseColor ? 1.f - GetValue(x, y) : GetValue(x, y);
            image.Set(x, y, TGAColor((std::uint8_t)(val * 255)));
        }
    return image;
}

TGAImage Buffer3f::GenerateImage() const
{
    pull();
    TGAImage image(m_Width, m_Height, TGAImage::RGB);
    for (int x = 0; x < m_Width; ++x)
        for (int y = 0; y < m_Height; ++y)
        {
            Vector3f c = GetValue(x, y);
            image.Set(x, y, TGAColor((std::uint8_t)(c.r * 254.99f), (std::uint8_t)(c.g * 254.99f),
                                     (std::uint8_t)(c.b * 254.99f)));
        }
    return image;
}

void Buffer3f::PaintColor(const Color3& color)
{
    if (m_Plane == FGL_PLANE_FRAME)
    {
        ForkerGL::ClearColor(color);
        return;
    }
    pull();
    for (size_t i = 0; i < (size_t)m_Width * m_Height; ++i)
        m_Host[i * 3] = color.x, m_Host[i * 3 + 1] = color.y, m_Host[i * 3 + 2] = color.z;
    m_HostDirty = true;
}

static void DeviceBlur(Buffer& b, int plane, int kind, bool& hostValid)
{
    if (plane < 0)
        throw std::runtime_error("Buffer blur: only ForkerGL's device planes can be post-processed "
                                 "(no CPU fallback)");
    ForkerGL::Check(fgl_blur(ForkerGL::Context(), plane, kind), "Buffer blur");
    hostValid = false;
    (void)b;
}

void Buffer1f::SimpleBlurDenoised() { push(); DeviceBlur(*this, m_Plane, FGL_BLUR_SIMPLE_3X3, m_HostValid); }
void Buffer1f::TwoPassGaussianBlurDenoised() { push(); DeviceBlur(*this, m_Plane, FGL_BLUR_TWO_PASS_GAUSSIAN, m_HostValid); }
void Buffer3f::SimpleBlurDenoised() { push(); DeviceBlur(*this, m_Plane, FGL_BLUR_SIMPLE_3X3, m_HostValid); }
void Buffer3f::TwoPassGaussianBlurDenoised() { push(); DeviceBlur(*this, m_Plane, FGL_BLUR_TWO_PASS_GAUSSIAN, m_HostValid); }
