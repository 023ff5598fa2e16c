/* ORACLE — test infrastructure only.  The reference decodes 8-bit textures with stb_image (framework/util/texture.cpp:106,
 * stbi_load); this shim compiles that header from the reference tree where it lies (3rdparty/stb/stb/stb_image.h, nothing is
 * copied) into oracle/_ref/libstb_ref.so so that tests/test_image_ldr.py can hold the host library's own JPEG / BMP / TGA /
 * PNG decoders to the reference's texel values.  Built only where the reference tree exists. */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_NO_STDIO
#include "stb_image.h"

/* returns malloc'ed w*h*channels bytes (free with stb_ref_free), channels as stored in the file (req_comp = 0, as the reference calls it) */
unsigned char *stb_ref_load(const unsigned char *file, int n, int *w, int *h, int *channels) { return stbi_load_from_memory(file, n, w, h, channels, 0); }
void stb_ref_free(void *p) { stbi_image_free(p); }
