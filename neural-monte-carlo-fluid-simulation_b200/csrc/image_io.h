// csrc/image_io.h -- the image files of the bindings' file-based entry points (bindings/zombie/demo/image.h):
// reading a source grid from a PFM file (Scene(config), scene.h:22-52) and writing the solution of bvc() as PFM or PNG plus
// the colour-mapped copy (demo/grid.h:9-33, colormap.h:14-35).  Host-only, no dependencies: the PNG encoder emits stored
// (uncompressed) deflate blocks.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace nmc_io {

inline bool hasExtension(const std::string& f, const char* ext) {
	const size_t p = f.find_last_of('.');
	if (p == std::string::npos) return false;
	std::string e = f.substr(p + 1);
	std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return (char)std::tolower(c); });
	return e == ext;
}

// Image<1>::readPFM (image.h:104-148): rows in file order (no flip on read); a 3-channel file is reduced to the
// grey value 0.299 r + 0.587 g + 0.114 b (setFromRGB, image.h:72-76), which a 1-channel file also goes through.
inline void readPfmGrey(const std::string& filename, int& h, int& w, std::vector<float>& out) {
	std::ifstream file(filename, std::ios::in | std::ios::binary);
	if (!file.is_open()) throw std::runtime_error("Error opening file: " + filename);
	char p = 0, type = 0;
	file >> p >> type;
	if (p != 'P' || (type != 'F' && type != 'f')) throw std::runtime_error("Invalid PFM file detected while reading " + filename);
	const int nc = type == 'F' ? 3 : 1;
	float scale = 0.0f;
	file >> w >> h >> scale;
	if (!file || w <= 0 || h <= 0) throw std::runtime_error("Invalid PFM header in " + filename);
	const uint16_t one = 1;
	const bool machineLittle = *reinterpret_cast<const uint8_t*>(&one) == 1;
	const bool flipBytes = (scale < 0) != machineLittle;
	file.ignore(1);
	std::vector<float> tmp((size_t)w*h*nc);
	file.read(reinterpret_cast<char*>(tmp.data()), (std::streamsize)(tmp.size()*sizeof(float)));
	if (!file) throw std::runtime_error("Truncated PFM file " + filename);
	if (flipBytes) for (float& v : tmp) { uint32_t u; memcpy(&u, &v, 4); u = (u >> 24) | ((u >> 8) & 0xFF00u) | ((u << 8) & 0xFF0000u) | (u << 24); memcpy(&v, &u, 4); }
	out.resize((size_t)w*h);
	for (size_t i = 0; i < out.size(); i++) {
		const float r = tmp[nc*i], g = tmp[nc*i + (nc == 3 ? 1 : 0)], b = tmp[nc*i + (nc == 3 ? 2 : 0)];
		out[i] = (float)(0.299*r + 0.587*g + 0.114*b);
	}
}

// Image<3>::writePFM (image.h:173-198): "PF", rows flipped (PFM stores the bottom row first)
inline void writePfm3(const std::string& filename, int h, int w, const std::vector<float>& rgb) {
	std::ofstream file(filename, std::ios::binary);
	if (!file) throw std::runtime_error("Error opening file: " + filename);
	file << "PF" << std::endl << w << " " << h << std::endl << "-1" << std::endl;
	std::vector<float> tmp((size_t)w*h*3);
	for (int i = 0; i < h; i++) memcpy(&tmp[(size_t)3*(h - i - 1)*w], &rgb[(size_t)3*i*w], (size_t)3*w*sizeof(float));
	file.write(reinterpret_cast<const char*>(tmp.data()), (std::streamsize)(tmp.size()*sizeof(float)));
}

inline uint32_t crc32(const uint8_t* d, size_t n, uint32_t crc = 0) {
	static uint32_t table[256];
	static bool init = false;
	if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
	crc = ~crc;
	for (size_t i = 0; i < n; i++) crc = table[(crc ^ d[i]) & 0xFF] ^ (crc >> 8);
	return ~crc;
}

// Image<3>::writePNG (image.h:200-215): 8 bits per channel, value = clamp(int(v * 255), 0, 255), rows in memory order
inline void writePng3(const std::string& filename, int h, int w, const std::vector<float>& rgb) {
	std::vector<uint8_t> raw((size_t)h*(1 + 3*(size_t)w));
	for (int i = 0; i < h; i++) {
		uint8_t* row = &raw[(size_t)i*(1 + 3*(size_t)w)];
		row[0] = 0; // filter: none
		for (int j = 0; j < 3*w; j++) row[1 + j] = (uint8_t)std::min(std::max((int)(rgb[(size_t)3*i*w + j]*255.0f), 0), 255);
	}
	std::vector<uint8_t> z;
	z.push_back(0x78); z.push_back(0x01);
	uint32_t a = 1, b = 0;
	for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
	for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
		const size_t len = std::min<size_t>(65535, raw.size() - off);
		z.push_back(off + len >= raw.size() ? 1 : 0);
		z.push_back((uint8_t)(len & 0xFF)); z.push_back((uint8_t)(len >> 8));
		z.push_back((uint8_t)(~len & 0xFF)); z.push_back((uint8_t)((~len >> 8) & 0xFF));
		z.insert(z.end(), raw.begin() + (std::ptrdiff_t)off, raw.begin() + (std::ptrdiff_t)(off + len));
		if (raw.empty()) break;
	}
	const uint32_t adler = (b << 16) | a;
	for (int k = 3; k >= 0; k--) z.push_back((uint8_t)(adler >> (8*k)));
	std::ofstream file(filename, std::ios::binary);
	if (!file) throw std::runtime_error("Failed to save image: " + filename);
	auto be32 = [](uint32_t v, uint8_t* o) { o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v; };
	auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
		std::vector<uint8_t> buf(8 + data.size() + 4);
		be32((uint32_t)data.size(), &buf[0]);
		memcpy(&buf[4], type, 4);
		if (!data.empty()) memcpy(&buf[8], data.data(), data.size());
		be32(crc32(&buf[4], 4 + data.size()), &buf[8 + data.size()]);
		file.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)buf.size());
	};
	const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	file.write(reinterpret_cast<const char*>(sig), 8);
	std::vector<uint8_t> ihdr(13);
	be32((uint32_t)w, &ihdr[0]); be32((uint32_t)h, &ihdr[4]);
	ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0; // 8-bit RGB
	chunk("IHDR", ihdr); chunk("IDAT", z); chunk("IEND", {});
}

inline void writeImage3(const std::string& filename, int h, int w, const std::vector<float>& rgb) { // Image::write (image.h:97-103)
	if (hasExtension(filename, "pfm")) writePfm3(filename, h, w, rgb); else writePng3(filename, h, w, rgb);
}

// applyColormap (colormap.h:14-22).  "turbo" (Google's public Turbo map) is carried as every 8th entry of the 256-entry
// look-up table plus the last one, joined by straight lines: within 0.007 of the table the reference indexes with
// int(value * 255).  Every other key falls back to grey, as unknown keys do in the reference.
inline void applyColormap(float v, const std::string& key, float* rgb) {
	v = std::min(std::max(v, 0.0f), 1.0f);
	if (key == "turbo") {
		static const float knots[33][3] = {
			{0.190f, 0.072f, 0.232f}, {0.225f, 0.164f, 0.451f}, {0.251f, 0.252f, 0.634f}, {0.268f, 0.338f, 0.780f}, {0.276f, 0.421f, 0.891f},
			{0.275f, 0.501f, 0.966f}, {0.259f, 0.580f, 0.999f}, {0.214f, 0.659f, 0.980f}, {0.158f, 0.736f, 0.923f}, {0.112f, 0.806f, 0.845f},
			{0.093f, 0.866f, 0.762f}, {0.120f, 0.912f, 0.687f}, {0.197f, 0.949f, 0.595f}, {0.305f, 0.977f, 0.490f}, {0.428f, 0.994f, 0.386f},
			{0.547f, 0.999f, 0.296f}, {0.644f, 0.990f, 0.234f}, {0.726f, 0.965f, 0.206f}, {0.805f, 0.925f, 0.205f}, {0.875f, 0.873f, 0.216f},
			{0.933f, 0.812f, 0.227f}, {0.973f, 0.747f, 0.225f}, {0.993f, 0.674f, 0.203f}, {0.996f, 0.587f, 0.169f}, {0.984f, 0.493f, 0.128f},
			{0.958f, 0.400f, 0.088f}, {0.921f, 0.315f, 0.055f}, {0.874f, 0.245f, 0.033f}, {0.816f, 0.185f, 0.018f}, {0.746f, 0.131f, 0.009f},
			{0.664f, 0.084f, 0.004f}, {0.571f, 0.045f, 0.005f}, {0.480f, 0.016f, 0.011f}};
		const int idx = (int)(v*255); // the table index the reference uses
		const int k = idx >= 248 ? 31 : idx/8;
		const int x0 = 8*k, x1 = k == 31 ? 255 : 8*k + 8;
		const float t = (float)(idx - x0)/(float)(x1 - x0);
		for (int c = 0; c < 3; c++) rgb[c] = knots[k][c] + t*(knots[k + 1][c] - knots[k][c]);
	} else rgb[0] = rgb[1] = rgb[2] = v;
}

} // namespace nmc_io
