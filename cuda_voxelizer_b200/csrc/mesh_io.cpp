// mesh_io.cpp — mesh loading for the CLI (stands in for trimesh2's TriMesh::read + need_faces + need_bbox,
// main.cpp:174-179).  Floats are parsed with strtof / std::from_chars, i.e. correctly rounded binary32 like trimesh2's
// sscanf("%f").  OBJ files — where ingest dominates once the voxelization takes a millisecond — are
// memory-mapped and parsed by all host threads at once (chunks cut at line ends, relative indices resolved after a
// prefix sum of the chunks' vertex counts); the line-by-line parser stays as the fallback and as the reference the
// parallel one is tested against (VOXCLI_SERIAL_LOADER=1).
#include <cerrno>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cli.h"

namespace voxcli {
namespace {

void finish_bbox(Mesh& m) {
	for (int k = 0; k < 3; k++) { m.bbox_min[k] = m.vertices[k]; m.bbox_max[k] = m.vertices[k]; }
	for (size_t i = 1; i < m.n_vertices(); i++)
		for (int k = 0; k < 3; k++) {
			const float v = m.vertices[3 * i + k];
			if (v < m.bbox_min[k]) m.bbox_min[k] = v;
			if (v > m.bbox_max[k]) m.bbox_max[k] = v;
		}
}

bool load_obj(const std::string& path, Mesh& m, std::string& error) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { error = "cannot open " + path; return false; }
	std::vector<char> line(1 << 16);
	std::vector<int32_t> poly;
	while (fgets(line.data(), (int)line.size(), f)) {
		const char* p = line.data();
		while (*p == ' ' || *p == '\t') p++;
		if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
			char* q = const_cast<char*>(p + 1);
			for (int k = 0; k < 3; k++) m.vertices.push_back(strtof(q, &q));
		} else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
			poly.clear();
			char* q = const_cast<char*>(p + 1);
			for (;;) {
				while (*q == ' ' || *q == '\t') q++;
				if (*q == '\0' || *q == '\n' || *q == '\r') break;
				char* e;
				long idx = strtol(q, &e, 10);
				if (e == q) break;
				const long nv = (long)m.n_vertices();
				poly.push_back((int32_t)(idx > 0 ? idx - 1 : nv + idx));
				q = e;
				while (*q && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') q++;   // skip /t/n
			}
			for (size_t k = 1; k + 1 < poly.size(); k++) { m.faces.push_back(poly[0]); m.faces.push_back(poly[k]); m.faces.push_back(poly[k + 1]); }
		}
	}
	fclose(f);
	return true;
}

// ---- parallel OBJ parser ----------------------------------------------------------------------
inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_eol(char c) { return c == '\n' || c == '\r'; }

// strtof's result on [p, end): correctly rounded binary32.  from_chars takes no leading '+' and no hex floats; anything it
// does not take goes through strtof on a bounded copy.  `p` advances past the number (or stays when there is none).
float parse_float(const char*& p, const char* end) {
	while (p < end && is_blank(*p)) p++;
	const char* q = p;
	if (q < end && *q == '+') q++;
	float v = 0.0f;
	const auto r = std::from_chars(q, end, v);
	if (r.ec == std::errc() && !(r.ptr < end && (*r.ptr == 'x' || *r.ptr == 'X'))) { p = r.ptr; return v; }
	char buf[128];
	size_t n = 0;
	while (p + n < end && n + 1 < sizeof(buf) && !is_eol(p[n])) { buf[n] = p[n]; n++; }
	buf[n] = '\0';
	char* e;
	v = strtof(buf, &e);
	p += e - buf;
	return v;
}

struct ObjChunk {
	std::vector<float> vertices;
	std::vector<int32_t> faces;              // 0-based, or (for relative indices, written <= 0) the written value minus one
	std::vector<std::pair<size_t, int32_t>> relative;      // {position in faces, vertices of this chunk seen before that face}
};

void parse_obj_chunk(const char* p, const char* end, ObjChunk& c) {
	std::vector<int32_t> poly;
	while (p < end) {
		const char* line_end = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
		if (!line_end) line_end = end;
		const char* q = p;
		while (q < line_end && is_blank(*q)) q++;
		if (q + 1 < line_end && q[0] == 'v' && is_blank(q[1])) {
			q++;
			for (int k = 0; k < 3; k++) c.vertices.push_back(parse_float(q, line_end));
		} else if (q + 1 < line_end && q[0] == 'f' && is_blank(q[1])) {
			q++;
			poly.clear();
			const int32_t seen = (int32_t)(c.vertices.size() / 3);
			for (;;) {
				while (q < line_end && is_blank(*q)) q++;
				if (q >= line_end || is_eol(*q)) break;
				const char* d = q;
				if (d < line_end && *d == '+') d++;
				long idx = 0;
				const auto r = std::from_chars(d, line_end, idx);
				if (r.ec != std::errc()) break;
				poly.push_back((int32_t)(idx - 1));      // idx <= 0 (stored < 0): relative to the vertices read so far, resolved by the caller
				q = r.ptr;
				while (q < line_end && !is_blank(*q) && !is_eol(*q)) q++;   // skip /t/n
			}
			for (size_t k = 1; k + 1 < poly.size(); k++)
				for (int32_t idx : {poly[0], poly[k], poly[k + 1]}) {
					if (idx < 0) c.relative.emplace_back(c.faces.size(), seen);
					c.faces.push_back(idx);
				}
		}
		p = line_end + 1;
	}
}

bool load_obj_parallel(const std::string& path, Mesh& m, std::string& error) {
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) { error = "cannot open " + path; return false; }
	struct stat st;
	if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return load_obj(path, m, error); }
	const size_t size = (size_t)st.st_size;
	void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (map == MAP_FAILED) return load_obj(path, m, error);
	const char* base = static_cast<const char*>(map);
	unsigned int n_threads = std::thread::hardware_concurrency();
	if (n_threads == 0) n_threads = 1;
	if (n_threads > 32) n_threads = 32;
	if (const char* env = getenv("VOXCLI_LOADER_THREADS")) { const int want = atoi(env); if (want >= 1 && want <= 256) n_threads = (unsigned int)want; }
	if (size < (size_t(1) << 20)) n_threads = 1;
	// chunk starts: byte offsets moved forward to the character after the next newline
	std::vector<size_t> cut(n_threads + 1, size);
	cut[0] = 0;
	for (unsigned int t = 1; t < n_threads; t++) {
		size_t at = size / n_threads * t;
		const char* nl = static_cast<const char*>(memchr(base + at, '\n', size - at));
		cut[t] = nl ? (size_t)(nl - base) + 1 : size;
		if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
	}
	std::vector<ObjChunk> chunks(n_threads);
	std::vector<std::thread> pool;
	for (unsigned int t = 1; t < n_threads; t++) pool.emplace_back([&, t] { parse_obj_chunk(base + cut[t], base + cut[t + 1], chunks[t]); });
	parse_obj_chunk(base + cut[0], base + cut[1], chunks[0]);
	for (auto& th : pool) th.join();
	munmap(map, size);
	size_t nv3 = 0, nf3 = 0;
	for (auto& c : chunks) { nv3 += c.vertices.size(); nf3 += c.faces.size(); }
	m.vertices.resize(nv3);
	m.faces.resize(nf3);
	size_t v_at = 0, f_at = 0;
	for (auto& c : chunks) {
		// relative (negative / zero) indices count back from the vertices read so far — in the whole file
		for (auto& rel : c.relative) c.faces[rel.first] = (int32_t)((long)(v_at / 3) + rel.second + c.faces[rel.first] + 1);
		if (!c.vertices.empty()) memcpy(m.vertices.data() + v_at, c.vertices.data(), c.vertices.size() * sizeof(float));
		if (!c.faces.empty()) memcpy(m.faces.data() + f_at, c.faces.data(), c.faces.size() * sizeof(int32_t));
		v_at += c.vertices.size();
		f_at += c.faces.size();
	}
	return true;
}

bool load_ply(const std::string& path, Mesh& m, std::string& error) {
	std::ifstream in(path, std::ios::binary);
	if (!in) { error = "cannot open " + path; return false; }
	std::string line, fmt;
	size_t nv = 0, nf = 0;
	std::vector<std::string> vprops;
	std::string list_count = "uchar", list_index = "int";
	int section = 0;
	if (!std::getline(in, line) || line.substr(0, 3) != "ply") { error = "not a PLY file"; return false; }
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		std::istringstream ss(line);
		std::string w;
		ss >> w;
		if (w == "format") ss >> fmt;
		else if (w == "element") { std::string name; size_t n; ss >> name >> n; if (name == "vertex") { nv = n; section = 1; } else if (name == "face") { nf = n; section = 2; } else section = 3; }
		else if (w == "property") {
			std::string t; ss >> t;
			if (section == 1) { std::string name; ss >> name; vprops.push_back(t + " " + name); }
			else if (section == 2 && t == "list") { ss >> list_count >> list_index; }
		} else if (w == "end_header") break;
	}
	auto tsize = [](const std::string& t) -> size_t {
		if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
		if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
		if (t == "double" || t == "float64") return 8;
		return 4;
	};
	m.vertices.resize(nv * 3);
	if (fmt == "ascii") {
		for (size_t i = 0; i < nv; i++) {
			std::getline(in, line);
			char* q = const_cast<char*>(line.c_str());
			for (size_t k = 0; k < vprops.size(); k++) { float v = strtof(q, &q); if (k < 3) m.vertices[3 * i + k] = v; }
		}
		for (size_t i = 0; i < nf; i++) {
			std::getline(in, line);
			char* q = const_cast<char*>(line.c_str());
			long n = strtol(q, &q, 10);
			std::vector<int32_t> poly;
			for (long k = 0; k < n; k++) poly.push_back((int32_t)strtol(q, &q, 10));
			for (size_t k = 1; k + 1 < poly.size(); k++) { m.faces.push_back(poly[0]); m.faces.push_back(poly[k]); m.faces.push_back(poly[k + 1]); }
		}
	} else if (fmt == "binary_little_endian") {
		size_t stride = 0;
		std::vector<size_t> off, sz;
		for (auto& p : vprops) { const std::string t = p.substr(0, p.find(' ')); off.push_back(stride); sz.push_back(tsize(t)); stride += tsize(t); }
		std::vector<char> rec(stride);
		for (size_t i = 0; i < nv; i++) {
			in.read(rec.data(), (std::streamsize)stride);
			for (int k = 0; k < 3 && k < (int)vprops.size(); k++) {
				if (sz[k] == 4) { float v; memcpy(&v, rec.data() + off[k], 4); m.vertices[3 * i + k] = v; }
				else if (sz[k] == 8) { double v; memcpy(&v, rec.data() + off[k], 8); m.vertices[3 * i + k] = (float)v; }
			}
		}
		const size_t cs = tsize(list_count), is = tsize(list_index);
		for (size_t i = 0; i < nf; i++) {
			unsigned long long n = 0;
			in.read(reinterpret_cast<char*>(&n), (std::streamsize)cs);
			std::vector<int32_t> poly;
			for (unsigned long long k = 0; k < n; k++) { long long idx = 0; in.read(reinterpret_cast<char*>(&idx), (std::streamsize)is); poly.push_back((int32_t)(is == 4 ? (int32_t)idx : idx)); }
			for (size_t k = 1; k + 1 < poly.size(); k++) { m.faces.push_back(poly[0]); m.faces.push_back(poly[k]); m.faces.push_back(poly[k + 1]); }
		}
	} else { error = "unsupported PLY format: " + fmt; return false; }
	return true;
}

}  // namespace

bool load_mesh(const std::string& path, Mesh& m, std::string& error) {
	m.vertices.clear();
	m.faces.clear();
	std::string ext = path.substr(path.find_last_of('.') == std::string::npos ? path.size() : path.find_last_of('.') + 1);
	for (auto& c : ext) c = (char)tolower(c);
	bool ok;
	const char* serial = getenv("VOXCLI_SERIAL_LOADER");
	if (ext == "obj") ok = (serial && serial[0] == '1') ? load_obj(path, m, error) : load_obj_parallel(path, m, error);
	else if (ext == "ply") ok = load_ply(path, m, error);
	else { error = "unsupported mesh format ." + ext + " (this build reads .obj and .ply; trimesh2 is not linked)"; return false; }
	if (!ok) return false;
	if (m.vertices.empty()) { error = "mesh has no vertices"; return false; }
	const int32_t nv = (int32_t)m.n_vertices();
	for (int32_t idx : m.faces)
		if (idx < 0 || idx >= nv) { error = "face index out of range"; return false; }
	finish_bbox(m);
	return true;
}

}  // namespace voxcli
