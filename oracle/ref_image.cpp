// ref_image.cpp -- the reference's image-texture load path on the host.
//
// TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/libref_stream.so).
// Restates the host half of RtwImage::Load (reference RtwImage.h:51-66) and
// FloatToByte (:100-105) on top of the reference's own stb translation unit
// (StbImageImpl.cpp, compiled where it lies): stbi_loadf gives linear floats
// (stb applies pow(x/255, 2.2)), which are re-quantised with (uchar)(256*f).
// The bytes this returns are what the reference uploads to the GPU (trap T8);
// tests/golden/make_golden.py stores them as the image fixture.
#include <cstring>

extern "C" {
float* stbi_loadf(char const* filename, int* x, int* y, int* channels_in_file, int desired_channels);
void stbi_image_free(void* retval_from_stbi_load);

// out may be NULL to query the size.  Returns 0 on success.
int ref_load_image_rgb8(const char* path, int* width, int* height, unsigned char* out, int capacity)
{
    int w = 0, h = 0, n = 0;
    float* f = stbi_loadf(path, &w, &h, &n, 3);
    if (!f) return -1;
    *width = w;
    *height = h;
    const int total = w * h * 3;
    if (out) {
        if (capacity < total) {
            stbi_image_free(f);
            return -2;
        }
        for (int k = 0; k < total; ++k) {
            const float v = f[k];
            out[k] = v <= 0.0f ? 0 : (1.0f <= v ? 255 : static_cast<unsigned char>(256.0f * v));
        }
    }
    stbi_image_free(f);
    return 0;
}
}
