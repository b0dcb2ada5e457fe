// tga.cpp -- ReadImageTGA (core/imageio.cpp:229-271) over the reference's bundled targa.c reader (tga_read_from_FILE,
// targa.c:165-279; tga_read_rle :282-330; tga_flip_horiz / tga_flip_vert :665-753; tga_color_unmap :761-797;
// tga_find_pixel :806-814), restated for the pixel formats whose bytes ReadImageTGA actually interprets: 8-bit grey
// (types 3, 11), 24 / 32-bit BGR(A) (types 2, 10), 8-bit indices into a 24 / 32-bit colour map (types 1, 9); raw or RLE,
// any of the four scan directions.  16-bit pixels or map entries are refused: ReadImageTGA reads three bytes per pixel
// whatever the depth, i.e. past such a pixel.
//
// Row order: ReadImageTGA normalises the image to left-to-right, top-to-bottom and then copies its rows from the LAST to
// the first ("origin in the lower left corner"): the returned image starts with the picture's bottom row.
#include "scene.h"

#include <cstdio>
#include <cstring>

namespace pb {

bool ReadImageTGA(const std::string &filename, int *xres, int *yres, std::vector<float> *rgb, std::string *err) {
    auto fail = [&](const std::string &why) {
        *err = "Unable to read from TGA file \"" + filename + "\" (" + why + ")";
        return false;
    };
    std::vector<uint8_t> d;
    {
        FILE *f = fopen(filename.c_str(), "rb");
        if (!f) return fail("error opening file");
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n < 0 || n > (1l << 30)) { fclose(f); return fail("file too large"); }
        d.resize((size_t)n);
        const bool ok = fread(d.data(), 1, d.size(), f) == d.size();
        fclose(f);
        if (!ok) return fail("premature end of file");
    }
    if (d.size() < 18) return fail("premature end of file");
    const uint8_t idLength = d[0], mapType = d[1], type = d[2];
    auto u16 = [&](size_t at) { return (uint32_t)d[at] | (uint32_t)d[at + 1] << 8; };
    const uint32_t mapOrigin = u16(3), mapLength = u16(5);
    const uint8_t mapDepth = d[7];
    const uint32_t width = u16(12), height = u16(14);
    const uint8_t depth = d[16], descriptor = d[17];
    if (mapType > 1) return fail("bad color map type");
    if (type == 0) return fail("no image data present");
    const bool mapped = type == 1 || type == 9, rle = type >= 9 && type <= 11, mono = type == 3 || type == 11;
    if (!(type >= 1 && type <= 3) && !rle) return fail("unknown image type");
    if (mapped && !mapType) return fail("image is color mapped but no color map is present");
    if (!mapped && mapType) return fail("non-color-mapped image has extraneous color map");
    if (mapType) {
        if (mapLength == 0) return fail("color map has zero length");
        if (mapDepth != 16 && mapDepth != 24 && mapDepth != 32) return fail("invalid color map depth");
        if (mapDepth == 16) return fail("16-bit color maps are not supported");
    }
    if (width == 0 || height == 0) return fail("the image dimensions are zero");
    if ((depth != 8 && depth != 16 && depth != 24 && depth != 32) || (depth != 8 && mapped)) return fail("invalid pixel depth");
    if (depth == 16 || (mono && depth != 8) || (!mono && !mapped && depth == 8)) return fail("pixel depth not supported for this image type");
    size_t p = 18 + (size_t)idLength;
    const size_t mapBpp = mapDepth / 8;
    const uint8_t *map = nullptr;
    if (mapType) {
        if (p + (size_t)mapLength * mapBpp > d.size()) return fail("premature end of file");
        map = d.data() + p;   // entry i of the file is colour index mapOrigin + i
        p += (size_t)mapLength * mapBpp;
    }
    const size_t bpp = depth / 8, nPixels = (size_t)width * height;
    if ((double)nPixels * (double)bpp > 130.0 * (double)d.size()) return fail("premature end of file");   // (RLE: at most 128 : 1)
    std::vector<uint8_t> pix(nPixels * bpp);
    if (!rle) {
        if (p + pix.size() > d.size()) return fail("premature end of file");
        memcpy(pix.data(), d.data() + p, pix.size());
    } else {
        size_t loaded = 0;
        uint8_t *pos = pix.data();
        while (loaded < nPixels) {
            if (p >= d.size()) return fail("premature end of file");
            const uint8_t b = d[p++];
            const size_t count = (size_t)(b & 0x7f) + 1;
            if (loaded + count > nPixels) return fail("RLE data is corrupt");
            if (b & 0x80) {
                if (p + bpp > d.size()) return fail("premature end of file");
                for (size_t i = 0; i < count; ++i, pos += bpp) memcpy(pos, d.data() + p, bpp);
                p += bpp;
            } else {
                if (p + bpp * count > d.size()) return fail("premature end of file");
                memcpy(pos, d.data() + p, bpp * count);
                p += bpp * count;
                pos += bpp * count;
            }
            loaded += count;
        }
    }
    const bool rightToLeft = (descriptor & 0x10) != 0, topToBottom = (descriptor & 0x20) != 0;
    rgb->resize(3 * nPixels);
    float *dst = rgb->data();
    for (uint32_t r = 0; r < height; ++r) {
        // output row r = picture row height-1-r (counted from the top) = file row r of a bottom-up file
        const uint32_t fileRow = topToBottom ? height - 1 - r : r;
        for (uint32_t x = 0; x < width; ++x) {
            const uint32_t fileX = rightToLeft ? width - 1 - x : x;
            const uint8_t *src = pix.data() + ((size_t)fileRow * width + fileX) * bpp;
            if (mapped) {
                const uint32_t index = *src;
                if (index >= mapOrigin + mapLength) return fail("color map index out of range");
                if (index < mapOrigin) return fail("color map index below the map's origin");   // (uninitialised memory in targa.c)
                src = map + (size_t)(index - mapOrigin) * mapBpp;
            }
            if (mono) {
                dst[0] = dst[1] = dst[2] = *src / 255.f;
            } else {   // BGR(A)
                dst[2] = src[0] / 255.f;
                dst[1] = src[1] / 255.f;
                dst[0] = src[2] / 255.f;
            }
            dst += 3;
        }
    }
    *xres = (int)width;
    *yres = (int)height;
    return true;
}

}  // namespace pb
