// texture.cpp — lazy upload of a Texture to the device.
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
