// Host object model: matrices, camera frame set-up, node transforms, BMP decode/encode, load-time
// gamma.  Load-time / per-frame host work only; nothing here runs per ray.
#include "rt.hpp"

#include <cstring>
#include <fstream>

namespace rt {

// ------------------------------------------------------------------ gfm:math 7.0.8 mat3d subset
// (the package is not vendored in the reference: dub.sdl:10, dub.selections.json:6; semantics as
// used at camera.d:102-104 and transform.d:24-50)
Matrix Matrix::identity() {
    Matrix m{};
    m.c[0][0] = m.c[1][1] = m.c[2][2] = 1.0;
    return m;
}
Matrix Matrix::scaledIdentity(double x, double y, double z) {
    Matrix m{};
    m.c[0][0] = x;
    m.c[1][1] = y;
    m.c[2][2] = z;
    return m;
}
static Matrix rotateAxis(int i, int j, double a) {
    Matrix m = Matrix::identity();
    const double ca = (double)cosl((long double)a), sa = (double)sinl((long double)a);
    m.c[i][i] = ca;
    m.c[i][j] = -sa;
    m.c[j][i] = sa;
    m.c[j][j] = ca;
    return m;
}
Matrix Matrix::rotateX(double a) { return rotateAxis(1, 2, a); }
Matrix Matrix::rotateY(double a) { return rotateAxis(2, 0, a); }
Matrix Matrix::rotateZ(double a) { return rotateAxis(0, 1, a); }
Matrix Matrix::operator*(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double sum = 0;
            for (int k = 0; k < 3; k++) sum += c[i][k] * o.c[k][j];
            r.c[i][j] = sum;
        }
    return r;
}
Matrix Matrix::inverse() const {
    const double det = c[0][0] * (c[1][1] * c[2][2] - c[2][1] * c[1][2]) - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                       c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
    const double invDet = 1 / det;
    Matrix r;
    r.c[0][0] = (c[1][1] * c[2][2] - c[2][1] * c[1][2]) * invDet;
    r.c[0][1] = -(c[0][1] * c[2][2] - c[0][2] * c[2][1]) * invDet;
    r.c[0][2] = (c[0][1] * c[1][2] - c[0][2] * c[1][1]) * invDet;
    r.c[1][0] = -(c[1][0] * c[2][2] - c[1][2] * c[2][0]) * invDet;
    r.c[1][1] = (c[0][0] * c[2][2] - c[0][2] * c[2][0]) * invDet;
    r.c[1][2] = -(c[0][0] * c[1][2] - c[1][0] * c[0][2]) * invDet;
    r.c[2][0] = (c[1][0] * c[2][1] - c[2][0] * c[1][1]) * invDet;
    r.c[2][1] = -(c[0][0] * c[2][1] - c[2][0] * c[0][1]) * invDet;
    r.c[2][2] = (c[0][0] * c[1][1] - c[1][0] * c[0][1]) * invDet;
    return r;
}
Matrix Matrix::transposed() const {
    Matrix r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.c[i][j] = c[j][i];
    return r;
}
Vector mul(const Vector& v, const Matrix& m) {
    return Vector(v.x * m.c[0][0] + v.y * m.c[1][0] + v.z * m.c[2][0], v.x * m.c[0][1] + v.y * m.c[1][1] + v.z * m.c[2][1],
                  v.x * m.c[0][2] + v.y * m.c[1][2] + v.z * m.c[2][2]);
}
double radians(double deg) { return (double)((long double)deg * (3.14159265358979323846264338327950288L / 180.0L)); }

// ------------------------------------------------------------------ Camera
void Camera::setFrameSize(uint32_t w, uint32_t h) {
    frameWidth = w;
    frameHeight = h;
    aspect = double(frameWidth) / double(frameHeight);
}

void Camera::beginFrame() {
    double x = -aspect, y = +1;
    const double lenXY = std::sqrt(0 + x * x + y * y + 0.0 * 0.0);  // |corner - center|
    const double wantedLength = (double)tanl((long double)radians(fov / 2));
    const double scaling = wantedLength / lenXY;
    x *= scaling;
    y *= scaling;
    const Matrix rotation = Matrix::rotateZ(radians(roll)) * Matrix::rotateX(radians(pitch)) * Matrix::rotateY(radians(yaw));
    upLeft_ = mul(Vector(x, y, 1), rotation);
    upRight_ = mul(Vector(-x, y, 1), rotation);
    downLeft_ = mul(Vector(x, -y, 1), rotation);
    rightDir_ = mul(Vector(1, 0, 0), rotation);
    upDir_ = mul(Vector(0, 1, 0), rotation);
    frontDir_ = mul(Vector(0, 0, 1), rotation);
    auto add = [&](Vector& v) { v.x += pos.x; v.y += pos.y; v.z += pos.z; };
    add(upLeft_);
    add(upRight_);
    add(downLeft_);
}

void Camera::move(double dx, double dy, double dz) {
    pos.x += dx * rightDir_.x; pos.y += dx * rightDir_.y; pos.z += dx * rightDir_.z;
    pos.x += dy * upDir_.x; pos.y += dy * upDir_.y; pos.z += dy * upDir_.z;
    pos.x += dz * frontDir_.x; pos.y += dz * frontDir_.y; pos.z += dz * frontDir_.z;
}

void Camera::rotate(double dYaw, double dRoll, double dPitch) {
    yaw += dYaw;
    roll += dRoll;
    pitch += dPitch;
    pitch = pitch < -90 ? -90 : (pitch > 90 ? 90 : pitch);
}

// ------------------------------------------------------------------ Transform
void Transform::reset() {
    transform = Matrix::identity();
    inverseTransform = transform.inverse();
    transposedInverse = inverseTransform.transposed();
    offset = Vector(0, 0, 0);
}
void Transform::scale(double x, double y, double z) {
    transform = transform * Matrix::scaledIdentity(x, y, z);
    inverseTransform = transform.inverse();
    transposedInverse = inverseTransform.transposed();
}
void Transform::rotate(double yaw, double pitch, double roll) {
    transform = transform * Matrix::rotateX(radians(pitch)) * Matrix::rotateY(radians(yaw)) * Matrix::rotateZ(radians(roll));
    inverseTransform = transform.inverse();
    transposedInverse = inverseTransform.transposed();
}
void Transform::translate(const Vector& v) { offset = v; }

// ------------------------------------------------------------------ BMP (imageio/bmp.d)
namespace {
struct Reader {
    const std::vector<uint8_t>& b;
    size_t pos = 0;
    void need(size_t n) const {
        if (pos + n > b.size()) throw ErrorLoadingImageException("BMP: unexpected end of file");
    }
    uint8_t u8() { need(1); return b[pos++]; }
    uint16_t u16() { need(2); uint16_t v = uint16_t(b[pos] | (b[pos + 1] << 8)); pos += 2; return v; }
    uint32_t u32() { need(4); uint32_t v = uint32_t(b[pos]) | (uint32_t(b[pos + 1]) << 8) | (uint32_t(b[pos + 2]) << 16) | (uint32_t(b[pos + 3]) << 24); pos += 4; return v; }
    void seek(size_t p) { pos = p; }
    void skip(size_t n) { pos += n; }
};
}  // namespace

Image<Color> loadBmpImage(const std::vector<uint8_t>& bytes) {
    Reader in{bytes};
    if (bytes.size() < 2 || bytes[0] != 'B' || bytes[1] != 'M') throw ImageIOException("Only files beginning with 'BM' are supported!");
    in.seek(10);
    const uint32_t pixelOffset = in.u32();
    const uint32_t dibSize = in.u32();
    int64_t width, height;
    uint32_t planes, bpp, colorsUsed = 0;
    const bool core = dibSize == 12;
    if (core) {
        width = (int16_t)in.u16();
        height = (int16_t)in.u16();
        planes = in.u16();
        bpp = in.u16();
    } else if (dibSize == 40 || dibSize == 52 || dibSize == 56 || dibSize == 108 || dibSize == 124) {
        width = (int32_t)in.u32();
        height = (int32_t)in.u32();
        planes = in.u16();
        bpp = in.u16();
        in.skip(16);  // compression, image size, x/y pixels per metre
        colorsUsed = in.u32();
    } else {
        throw ErrorLoadingImageException("BMP: unsupported DIB header (" + std::to_string(dibSize) + " bytes)");
    }
    if (planes != 1) throw ErrorLoadingImageException("Only .bmp files with 1 color plane are supported. Not: " + std::to_string(planes));
    switch (bpp) {
        case 1: case 2: case 4: case 8: case 16: case 24: case 32: case 64: break;
        default: throw ErrorLoadingImageException("Only .bmp files with 1, 2, 4, 8, 16, 24, 32 or 64 bpp are supported. Not: " + std::to_string(bpp));
    }
    if (width <= 0 || height <= 0) throw ErrorLoadingImageException("BMP: non-positive image size");

    std::vector<uint32_t> palette;
    if (bpp <= 8) {
        const uint32_t count = core ? (1u << bpp) : (colorsUsed ? colorsUsed : (1u << bpp));
        in.seek(14 + dibSize);
        for (uint32_t i = 0; i < count; i++) {
            uint32_t bl = in.u8(), gr = in.u8(), re = in.u8(), al = core ? 0 : in.u8();
            palette.push_back(bl | (gr << 8) | (re << 16) | (al << 24));
        }
    }
    Image<Color> img((size_t)width, (size_t)height);
    in.seek(pixelOffset);
    const size_t W = (size_t)width, H = (size_t)height;
    const size_t rowBytes = bpp / 8 * W, rowStride = ((bpp * W + 31) / 32) * 4;
    if (bpp == 24 || bpp == 32) {
        for (size_t y = H; y-- > 0;) {  // scanlines are stored bottom-up
            for (size_t x = 0; x < W; x++) {
                uint32_t bl = in.u8(), gr = in.u8(), re = in.u8();
                if (bpp == 32) in.u8();
                img(x, y) = Color(bl | (gr << 8) | (re << 16));
            }
            in.skip(rowStride - rowBytes);
        }
    } else if (bpp <= 8) {
        // the reference consumes exactly `width` bytes per scanline and unpacks 8/bpp pixels from each
        const size_t perByte = 8 / bpp;
        const uint32_t mask = (1u << bpp) - 1;
        for (size_t y = H; y-- > 0;) {
            for (size_t i = 0; i < W; i++) {
                const uint8_t pack = in.u8();
                for (size_t k = 0; k < perByte; k++) {
                    const uint32_t idx = (pack >> (bpp * (perByte - 1 - k))) & mask;
                    const size_t x = i * perByte + k;
                    if (x >= W || idx >= palette.size()) throw ErrorLoadingImageException("BMP: index out of range while unpacking");
                    img(x, y) = Color(palette[idx]);
                }
            }
        }
    } else {
        throw ErrorLoadingImageException("Not implemented: bpp > 8 && bpp != 24 && bpp != 32");
    }
    return img;
}

// 24-bpp BITMAPINFOHEADER writer (imageio/bmp.d:195-237): fileSize = 14 + 40 + 3 W H, sizeOfPixelArray = fileSize - 54,
// 72.dpiToPPM = lrint(72 * 100 / 2.54) = 2835 (bmp.d:253), rows from the bottom (foreach_reverse), each pixel the low three
// little-endian bytes of Color.toRGB32 (b, g, r), rows NOT padded (see rt.hpp).
std::vector<uint8_t> saveBmp(const Image<uint32_t>& img, bool padRows) {
    const size_t W = img.width, H = img.height;
    const size_t stride = padRows ? (W * 3 + 3) / 4 * 4 : W * 3;
    std::vector<uint8_t> out(54 + stride * H, 0);
    auto w32 = [&](size_t o, uint32_t v) { out[o] = v & 0xff; out[o + 1] = (v >> 8) & 0xff; out[o + 2] = (v >> 16) & 0xff; out[o + 3] = (v >> 24) & 0xff; };
    auto w16 = [&](size_t o, uint16_t v) { out[o] = v & 0xff; out[o + 1] = (v >> 8) & 0xff; };
    out[0] = 'B'; out[1] = 'M';
    w32(2, (uint32_t)out.size());
    w32(10, 54);
    w32(14, 40);
    w32(18, (uint32_t)W);
    w32(22, (uint32_t)H);
    w16(26, 1);
    w16(28, 24);
    w32(34, (uint32_t)(stride * H));
    w32(38, 2835);
    w32(42, 2835);
    for (size_t y = 0; y < H; y++) {
        uint8_t* row = &out[54 + stride * (H - 1 - y)];
        for (size_t x = 0; x < W; x++) {
            const uint32_t v = img(x, y);
            row[3 * x + 0] = v & 0xff;
            row[3 * x + 1] = (v >> 8) & 0xff;
            row[3 * x + 2] = (v >> 16) & 0xff;
        }
    }
    return out;
}

// ------------------------------------------------------------------ Bitmap (rt/bitmap.d)
void Bitmap::loadImage(const std::string& filename) {
    std::string ext;
    const size_t dot = filename.find_last_of('.');
    if (dot != std::string::npos) ext = filename.substr(dot);
    for (auto& ch : ext) ch = (char)tolower((unsigned char)ch);
    if (ext == ".exr") throw ImageIOException("Not implemented");  // bitmap.d:170-173
    if (ext != ".bmp") throw UnknownImageTypeException();
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw ImageIOException("cannot read '" + filename + "'");
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    data = loadBmpImage(bytes);
}

void Bitmap::decompressGamma_sRGB() {  // bitmap.d:116-126
    auto fn = [](float x) -> float {
        if (x == 0) return 0.0f;
        if (x == 1) return 1.0f;
        if (x <= 0.04045f) return x / 12.92f;
        return (float)powl((long double)((x + 0.055f) / 1.055f), (long double)2.4f);
    };
    for (auto& p : data.pixels) { p.r = fn(p.r); p.g = fn(p.g); p.b = fn(p.b); }
}

void Bitmap::decompressGamma(float gamma) {  // bitmap.d:129-136
    auto fn = [gamma](float x) -> float {
        if (x == 0) return 0.0f;
        if (x == 1) return 1.0f;
        return (float)powl((long double)x, (long double)gamma);
    };
    for (auto& p : data.pixels) { p.r = fn(p.r); p.g = fn(p.g); p.b = fn(p.b); }
}

}  // namespace rt
