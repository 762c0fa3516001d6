// io.hpp — host-side file formats of the reference, kept byte/parse compatible:
//   P3 PPM writer  = Screen::write_to_ppm / write_to_ppm_file   (src/screen.rs:40-59)
//   P3 PPM reader  = Screen::from_ppm_p3                         (src/screen.rs:61-95)
//   ASCII PLY      = TriangleModel::load_from_file               (src/model.rs:13-62)
// plus binary fast paths the reference does not have: P6 PPM out/in, binary_little_endian PLY in.
// The text formats stay byte / parse compatible; the loaders pick the binary path from the file's own header.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rtb {

// rows emitted top-down (j = H-1 .. 0), one "r g b\n" line per pixel; integer-valued doubles print
// without a decimal point (Rust `{}` on f64)
bool write_ppm_p3(const char* path_or_null, const double* screen, int32_t width, int32_t height, std::string& err);

// texels returned in file order (row 0 = first row of the file), 3 doubles per pixel
bool read_ppm_p3(const char* path, int32_t& width, int32_t& height, std::vector<double>& rgb, std::string& err);

// P6: same header, then W*H*3 bytes, rows top-down; channel values are the Screen's integers 0..255
bool write_ppm_p6(const char* path_or_null, const double* screen, int32_t width, int32_t height, std::string& err);
// magic "P6" -> binary samples (maxval <= 255), anything else -> the P3 parser above
bool read_ppm_any(const char* path, int32_t& width, int32_t& height, std::vector<double>& rgb, std::string& err);

// header scan for "element vertex N" / "element face M" / "end_header"; vertex = first three tokens
// times `scale`; face = tokens 1..3 (token 0, the count, is ignored)
bool read_ply_ascii(const char* path, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err);
// "format binary_little_endian 1.0" -> fixed-stride records (vertex = its first three float/double properties,
// face = one list property, triangles only); "format ascii" -> read_ply_ascii
bool read_ply_any(const char* path, double scale, std::vector<double>& verts, std::vector<uint32_t>& faces, std::string& err);
// writer used by tools and tests: vertex = 3 float32, face = uchar 3 + 3 int32
bool write_ply_binary(const char* path, const std::vector<double>& verts, const std::vector<uint32_t>& faces, std::string& err);

} // namespace rtb
