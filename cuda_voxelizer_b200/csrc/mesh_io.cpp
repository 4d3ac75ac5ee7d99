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
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "cli.h"

namespace voxcli {
namespace {

// trimesh2's tess() (TriMesh_io.cc, tag 2022.03.04 — not vendored in the reference tree, restated from its documented behaviour):
// a triangle as is; a QUAD along its SHORTER diagonal (dist2(p0,p2) < dist2(p1,p3) ? from corner 0 : from corner 1); five and
// more corners as a fan from corner 0.  Quads need vertex positions, so loaders that parse in parallel record them and
// split them once all vertices are known (split_quads).
inline void push_tri(std::vector<int32_t>& faces, int32_t a, int32_t b, int32_t c) { faces.push_back(a); faces.push_back(b); faces.push_back(c); }
inline float dist2(const std::vector<float>& v, int32_t a, int32_t b) {
	const float dx = v[3 * (size_t)a] - v[3 * (size_t)b], dy = v[3 * (size_t)a + 1] - v[3 * (size_t)b + 1], dz = v[3 * (size_t)a + 2] - v[3 * (size_t)b + 2];
	return dx * dx + dy * dy + dz * dz;
}
inline void quad_split(const std::vector<float>& verts, const int32_t q[4], int32_t out[6]) {
	const size_t nv = verts.size() / 3;
	int i = 1;
	bool ok = true;
	for (int k = 0; k < 4; k++) ok = ok && q[k] >= 0 && (size_t)q[k] < nv;
	if (ok) i = dist2(verts, q[0], q[2]) < dist2(verts, q[1], q[3]) ? 0 : 1;         // (bad indices are reported by the range check later)
	out[0] = q[i]; out[1] = q[(i + 1) % 4]; out[2] = q[(i + 2) % 4];
	out[3] = q[i]; out[4] = q[(i + 2) % 4]; out[5] = q[(i + 3) % 4];
}
// Serial loaders: vertices of the polygon are known already.
inline void tess(const std::vector<float>& verts, const std::vector<int32_t>& poly, std::vector<int32_t>& faces) {
	if (poly.size() < 3) return;
	if (poly.size() == 4) {
		int32_t t[6];
		quad_split(verts, poly.data(), t);
		push_tri(faces, t[0], t[1], t[2]); push_tri(faces, t[3], t[4], t[5]);
		return;
	}
	for (size_t k = 1; k + 1 < poly.size(); k++) push_tri(faces, poly[0], poly[k], poly[k + 1]);
}

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
			tess(m.vertices, poly, m.faces);
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

// A decimal integer on [p, end) (indices and list counts: a float would lose integers above 2^24); 0 when there is none.
long parse_long(const char*& p, const char* end) {
	while (p < end && is_blank(*p)) p++;
	const char* q = p;
	if (q < end && *q == '+') q++;
	long v = 0;
	const auto r = std::from_chars(q, end, v);
	if (r.ec != std::errc()) return 0;
	p = r.ptr;
	// "3.0"-style tokens: skip a fractional part
	if (p < end && *p == '.') { p++; while (p < end && *p >= '0' && *p <= '9') p++; }
	return v;
}

struct ObjChunk {
	std::vector<float> vertices;
	std::vector<int32_t> faces;              // 0-based, or (for relative indices, written <= 0) the written value minus one
	std::vector<std::pair<size_t, int32_t>> relative;      // {position in faces, vertices of this chunk seen before that face}
	std::vector<size_t> quads;               // positions in faces of quads, stored as corner 0,1,2, 0,2,3 until the vertices are known
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
			if (poly.size() == 4) c.quads.push_back(c.faces.size());
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
	std::vector<size_t> quads;
	for (auto& c : chunks) {
		// relative (negative / zero) indices count back from the vertices read so far — in the whole file
		for (auto& rel : c.relative) c.faces[rel.first] = (int32_t)((long)(v_at / 3) + rel.second + c.faces[rel.first] + 1);
		if (!c.vertices.empty()) memcpy(m.vertices.data() + v_at, c.vertices.data(), c.vertices.size() * sizeof(float));
		if (!c.faces.empty()) memcpy(m.faces.data() + f_at, c.faces.data(), c.faces.size() * sizeof(int32_t));
		for (size_t q : c.quads) quads.push_back(f_at + q);
		v_at += c.vertices.size();
		f_at += c.faces.size();
	}
	// quads along their shorter diagonal, now that every vertex is known (stored as 0,1,2, 0,2,3)
	for (size_t at : quads) {
		const int32_t q[4] = {m.faces[at], m.faces[at + 1], m.faces[at + 2], m.faces[at + 5]};
		quad_split(m.vertices, q, m.faces.data() + at);
	}
	return true;
}

// ---- PLY ------------------------------------------------------------------------------------------
// ASCII, binary little- and big-endian; vertex properties x, y, z wherever they sit in the record (float or double);
// faces as a list property (polygons through tess()) and/or triangle strips (element tristrips, -1 restarts a strip, every
// second triangle flipped — what trimesh2's need_faces() unpacks).  The file is memory-mapped.  The common case — a binary file
// whose faces are all triangles with a one-byte count and four-byte indices — is copied out by all host threads (each checks the
// count bytes of its range first); ASCII files are cut at line ends and parsed by all threads like OBJ files.
struct PlyProp { std::string type, name; bool list = false; std::string count_type; };
struct PlyElement { std::string name; size_t count = 0; std::vector<PlyProp> props; };

size_t ply_size(const std::string& t) {
	if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
	if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
	if (t == "double" || t == "float64") return 8;
	return 4;            // int, uint, float, int32, uint32, float32
}
bool ply_is_float(const std::string& t) { return t == "float" || t == "float32" || t == "double" || t == "float64"; }
bool ply_is_signed(const std::string& t) { return t == "char" || t == "int8" || t == "short" || t == "int16" || t == "int" || t == "int32"; }

// one scalar of type t at p (byte-swapped when the file's endianness is not the host's), as double
double ply_scalar(const char* p, const std::string& t, bool swap) {
	const size_t n = ply_size(t);
	unsigned char b[8];
	for (size_t k = 0; k < n; k++) b[k] = (unsigned char)p[swap ? n - 1 - k : k];
	if (ply_is_float(t)) {
		if (n == 4) { float v; memcpy(&v, b, 4); return v; }
		double v; memcpy(&v, b, 8); return v;
	}
	unsigned long long u = 0;
	memcpy(&u, b, n);                                  // host is little-endian (x86-64)
	if (ply_is_signed(t)) {
		if (n == 1) return (signed char)u;
		if (n == 2) return (short)u;
		return (int)u;
	}
	return (double)u;
}

unsigned int loader_threads(size_t bytes) {
	unsigned int n = std::thread::hardware_concurrency();
	if (n == 0) n = 1;
	if (n > 32) n = 32;
	if (const char* env = getenv("VOXCLI_LOADER_THREADS")) { const int want = atoi(env); if (want >= 1 && want <= 256) n = (unsigned int)want; }
	if (bytes < (size_t(1) << 20)) n = 1;
	return n;
}
template <typename F>
void run_threads(unsigned int n, F fn) {
	std::vector<std::thread> pool;
	for (unsigned int t = 1; t < n; t++) pool.emplace_back([&fn, t] { fn(t); });
	fn(0u);
	for (auto& th : pool) th.join();
}

void unpack_strip(const std::vector<int32_t>& s, std::vector<int32_t>& faces) {
	size_t start = 0;
	for (size_t i = 0; i <= s.size(); i++) {
		if (i < s.size() && s[i] != -1) continue;
		for (size_t k = start + 2; k < i; k++) {
			if ((k - start) % 2 == 0) push_tri(faces, s[k - 2], s[k - 1], s[k]);
			else push_tri(faces, s[k - 1], s[k - 2], s[k]);
		}
		start = i + 1;
	}
}

bool load_ply(const std::string& path, Mesh& m, std::string& error) {
	const int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) { error = "cannot open " + path; return false; }
	struct stat st;
	if (fstat(fd, &st) != 0 || st.st_size < 4) { close(fd); error = "not a PLY file"; return false; }
	const size_t size = (size_t)st.st_size;
	void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (map == MAP_FAILED) { error = "cannot map " + path; return false; }
	const char* base = static_cast<const char*>(map);
	const char* end = base + size;
	struct Unmap { void* p; size_t n; ~Unmap() { munmap(p, n); } } unmap{map, size};
	// header
	const char* p = base;
	auto next_line = [&](std::string& line) -> bool {
		if (p >= end) return false;
		const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
		const char* le = nl ? nl : end;
		line.assign(p, le);
		if (!line.empty() && line.back() == '\r') line.pop_back();
		p = nl ? nl + 1 : end;
		return true;
	};
	std::string line, fmt;
	if (!next_line(line) || line.substr(0, 3) != "ply") { error = "not a PLY file"; return false; }
	std::vector<PlyElement> elements;
	bool header_done = false;
	while (next_line(line)) {
		std::istringstream ss(line);
		std::string w;
		ss >> w;
		if (w == "format") ss >> fmt;
		else if (w == "element") { PlyElement e; ss >> e.name >> e.count; elements.push_back(e); }
		else if (w == "property" && !elements.empty()) {
			PlyProp pr;
			ss >> pr.type;
			if (pr.type == "list") { pr.list = true; ss >> pr.count_type >> pr.type; }
			ss >> pr.name;
			elements.back().props.push_back(pr);
		} else if (w == "end_header") { header_done = true; break; }
	}
	if (!header_done) { error = "PLY header has no end_header"; return false; }
	const bool ascii = fmt == "ascii";
	if (!ascii && fmt != "binary_little_endian" && fmt != "binary_big_endian") { error = "unsupported PLY format: " + fmt; return false; }
	const bool swap = fmt == "binary_big_endian";
	std::vector<int32_t> poly;
	for (const PlyElement& e : elements) {
		const bool is_vertex = e.name == "vertex", is_face = e.name == "face", is_strips = e.name == "tristrips";
		int ix = -1, iy = -1, iz = -1, ilist = -1;
		bool any_list = false;
		for (size_t k = 0; k < e.props.size(); k++) {
			if (e.props[k].list) { any_list = true; if (ilist < 0) ilist = (int)k; }
			if (e.props[k].name == "x") ix = (int)k; else if (e.props[k].name == "y") iy = (int)k; else if (e.props[k].name == "z") iz = (int)k;
			if (e.props[k].list && (e.props[k].name == "vertex_indices" || e.props[k].name == "vertex_index")) ilist = (int)k;
		}
		if (is_vertex && (ix < 0 || iy < 0 || iz < 0)) { error = "PLY vertex element has no x / y / z properties"; return false; }
		if ((is_face || is_strips) && ilist < 0 && e.count) { error = "PLY " + e.name + " element has no index list"; return false; }
		if (is_vertex) m.vertices.resize(e.count * 3);
		if (ascii) {
			// the element's lines: count them out, then cut the range at line ends for the threads
			const char* first = p;
			const char* q = p;
			for (size_t i = 0; i < e.count; i++) {
				const char* nl = q < end ? static_cast<const char*>(memchr(q, '\n', (size_t)(end - q))) : nullptr;
				if (!nl) { if (q < end && i + 1 == e.count) { q = end; break; } error = "PLY file ends inside element " + e.name; return false; }
				q = nl + 1;
			}
			p = q;
			if (!(is_vertex || is_face || is_strips)) continue;
			const size_t bytes = (size_t)(q - first);
			const unsigned int T = loader_threads(bytes);
			std::vector<const char*> cut(T + 1, q);
			cut[0] = first;
			for (unsigned int t = 1; t < T; t++) {
				const char* at = first + bytes / T * t;
				const char* nl = static_cast<const char*>(memchr(at, '\n', (size_t)(q - at)));
				cut[t] = nl ? nl + 1 : q;
				if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
			}
			std::vector<size_t> lines(T + 1, 0);
			run_threads(T, [&](unsigned int t) { size_t n = 0; for (const char* c = cut[t]; c < cut[t + 1]; c++) n += *c == '\n'; lines[t + 1] = n; });
			if (q == end && size && end[-1] != '\n') lines[T]++;         // a last line without a newline
			for (unsigned int t = 0; t < T; t++) lines[t + 1] += lines[t];
			std::vector<std::vector<int32_t>> part(T);
			std::vector<std::vector<size_t>> part_quads(T);
			run_threads(T, [&](unsigned int t) {
				const char* c = cut[t];
				std::vector<int32_t> pl;
				for (size_t row = lines[t]; row < lines[t + 1] && row < e.count; row++) {
					const char* nl = static_cast<const char*>(memchr(c, '\n', (size_t)(cut[t + 1] - c)));
					const char* le = nl ? nl : cut[t + 1];
					const char* r = c;
					if (is_vertex) {
						for (int k = 0; k < (int)e.props.size(); k++) {
							if (e.props[k].list) { const long n = parse_long(r, le); for (long j = 0; j < n; j++) parse_float(r, le); continue; }
							const float v = parse_float(r, le);
							if (k == ix) m.vertices[3 * row] = v; else if (k == iy) m.vertices[3 * row + 1] = v; else if (k == iz) m.vertices[3 * row + 2] = v;
						}
					} else {
						for (int k = 0; k < (int)e.props.size(); k++) {
							if (!e.props[k].list) { parse_float(r, le); continue; }
							const long n = parse_long(r, le);
							pl.clear();
							for (long j = 0; j < n; j++) { const int32_t idx = (int32_t)parse_long(r, le); if (k == ilist) pl.push_back(idx); }
							if (k != ilist) continue;
							if (is_strips) unpack_strip(pl, part[t]);
							else if (pl.size() == 4) { part_quads[t].push_back(part[t].size()); push_tri(part[t], pl[0], pl[1], pl[2]); push_tri(part[t], pl[0], pl[2], pl[3]); }
							else for (size_t j = 1; j + 1 < pl.size(); j++) push_tri(part[t], pl[0], pl[j], pl[j + 1]);
						}
					}
					c = nl ? nl + 1 : cut[t + 1];
				}
			});
			for (unsigned int t = 0; t < T; t++) {
				const size_t at = m.faces.size();
				m.faces.insert(m.faces.end(), part[t].begin(), part[t].end());
				for (size_t qd : part_quads[t]) {
					const int32_t qv[4] = {m.faces[at + qd], m.faces[at + qd + 1], m.faces[at + qd + 2], m.faces[at + qd + 5]};
					quad_split(m.vertices, qv, m.faces.data() + at + qd);        // (vertex element precedes the faces in every PLY writer)
				}
			}
			continue;
		}
		// binary
		if (!any_list) {
			size_t stride = 0;
			std::vector<size_t> off;
			for (auto& pr : e.props) { off.push_back(stride); stride += ply_size(pr.type); }
			if ((size_t)(end - p) < stride * e.count) { error = "PLY file ends inside element " + e.name; return false; }
			if (is_vertex) {
				const char* rec0 = p;
				const unsigned int T = loader_threads(stride * e.count);
				run_threads(T, [&](unsigned int t) {
					for (size_t i = e.count * t / T; i < e.count * (t + 1) / T; i++) {
						const char* rec = rec0 + i * stride;
						m.vertices[3 * i] = (float)ply_scalar(rec + off[ix], e.props[ix].type, swap);
						m.vertices[3 * i + 1] = (float)ply_scalar(rec + off[iy], e.props[iy].type, swap);
						m.vertices[3 * i + 2] = (float)ply_scalar(rec + off[iz], e.props[iz].type, swap);
					}
				});
			}
			p += stride * e.count;
			continue;
		}
		// records with lists.  Fast path: faces whose only property is a list with a 1-byte count and 4-byte indices, all of them triangles
		if (is_face && e.props.size() == 1 && ply_size(e.props[0].count_type) == 1 && ply_size(e.props[0].type) == 4 && !swap &&
		    (size_t)(end - p) >= e.count * 13) {
			const char* rec0 = p;
			const unsigned int T = loader_threads(e.count * 13);
			std::vector<char> all3(T, 1);
			run_threads(T, [&](unsigned int t) { for (size_t i = e.count * t / T; i < e.count * (t + 1) / T; i++) if (rec0[i * 13] != 3) { all3[t] = 0; break; } });
			bool ok = true;
			for (char c : all3) ok = ok && c;
			if (ok) {
				const size_t at = m.faces.size();
				m.faces.resize(at + e.count * 3);
				run_threads(T, [&](unsigned int t) { for (size_t i = e.count * t / T; i < e.count * (t + 1) / T; i++) memcpy(m.faces.data() + at + 3 * i, rec0 + i * 13 + 1, 12); });
				p += e.count * 13;
				continue;
			}
		}
		for (size_t i = 0; i < e.count; i++) {
			for (int k = 0; k < (int)e.props.size(); k++) {
				const PlyProp& pr = e.props[k];
				if (!pr.list) {
					if ((size_t)(end - p) < ply_size(pr.type)) { error = "PLY file ends inside element " + e.name; return false; }
					if (is_vertex && (k == ix || k == iy || k == iz)) m.vertices[3 * i + (k == ix ? 0 : k == iy ? 1 : 2)] = (float)ply_scalar(p, pr.type, swap);
					p += ply_size(pr.type);
					continue;
				}
				if ((size_t)(end - p) < ply_size(pr.count_type)) { error = "PLY file ends inside element " + e.name; return false; }
				const size_t n = (size_t)ply_scalar(p, pr.count_type, swap);
				p += ply_size(pr.count_type);
				const size_t is = ply_size(pr.type);
				if ((size_t)(end - p) < n * is) { error = "PLY file ends inside element " + e.name; return false; }
				if ((is_face || is_strips) && k == ilist) {
					poly.clear();
					for (size_t j = 0; j < n; j++) poly.push_back((int32_t)ply_scalar(p + j * is, pr.type, swap));
					if (is_strips) unpack_strip(poly, m.faces); else tess(m.vertices, poly, m.faces);
				}
				p += n * is;
			}
		}
	}
	return true;
}

// ---- OFF, STL, 3DS ------------------------------------------------------------------------------
// The other formats the reference's help text names or trimesh2 reads for it (main.cpp:121 ".ply, .obj, .3ds"): small, serial
// readers over the mapped file.  OFF: "OFF", counts, vertices, polygons (through tess()).  STL: binary (80-byte header, count,
// 50-byte records) or ASCII ("vertex x y z" lines); three fresh vertices per facet, as trimesh2 reads it (no welding).
// 3DS: chunk tree 0x4D4D > 0x3D3D > 0x4000 (named object) > 0x4100 (mesh) > 0x4110 vertices / 0x4120 faces; objects are appended
// with their vertex offsets.
struct Mapped {
	const char* base = nullptr; size_t size = 0;
	bool open_file(const std::string& path, std::string& error) {
		const int fd = open(path.c_str(), O_RDONLY);
		if (fd < 0) { error = "cannot open " + path; return false; }
		struct stat st;
		if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); error = "empty file " + path; return false; }
		size = (size_t)st.st_size;
		void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
		close(fd);
		if (map == MAP_FAILED) { error = "cannot map " + path; return false; }
		base = static_cast<const char*>(map);
		return true;
	}
	~Mapped() { if (base) munmap(const_cast<char*>(base), size); }
};

// next whitespace-separated token on [p, end), skipping '#' comments to the end of their line; false at the end
bool next_token(const char*& p, const char* end, const char*& tok, const char*& tok_end) {
	for (;;) {
		while (p < end && (is_blank(*p) || is_eol(*p))) p++;
		if (p < end && *p == '#') { while (p < end && *p != '\n') p++; continue; }
		break;
	}
	if (p >= end) return false;
	tok = p;
	while (p < end && !is_blank(*p) && !is_eol(*p)) p++;
	tok_end = p;
	return true;
}

bool load_off(const std::string& path, Mesh& m, std::string& error) {
	Mapped f;
	if (!f.open_file(path, error)) return false;
	const char* p = f.base; const char* end = f.base + f.size;
	const char *t, *te;
	if (!next_token(p, end, t, te)) { error = "empty OFF file"; return false; }
	// the "OFF" keyword may be missing, or be followed by the counts on the same line
	if (!(te - t >= 3 && (memcmp(te - 3, "OFF", 3) == 0))) p = t;
	long counts[3] = {0, 0, 0};
	for (int k = 0; k < 3; k++) {
		if (!next_token(p, end, t, te)) { error = "OFF header is incomplete"; return false; }
		const char* q = t; counts[k] = parse_long(q, te);
	}
	if (counts[0] <= 0 || counts[1] < 0) { error = "OFF header has no vertices"; return false; }
	m.vertices.resize((size_t)counts[0] * 3);
	for (size_t i = 0; i < m.vertices.size(); i++) {
		if (!next_token(p, end, t, te)) { error = "OFF file ends inside the vertices"; return false; }
		const char* q = t; m.vertices[i] = parse_float(q, te);
	}
	std::vector<int32_t> poly;
	for (long i = 0; i < counts[1]; i++) {
		if (!next_token(p, end, t, te)) { error = "OFF file ends inside the faces"; return false; }
		const char* q = t; const long n = parse_long(q, te);
		poly.clear();
		for (long k = 0; k < n; k++) {
			if (!next_token(p, end, t, te)) { error = "OFF file ends inside the faces"; return false; }
			q = t; poly.push_back((int32_t)parse_long(q, te));
		}
		tess(m.vertices, poly, m.faces);
		while (p < end && !is_eol(*p)) p++;            // colours behind the indices
	}
	return true;
}

bool load_stl(const std::string& path, Mesh& m, std::string& error) {
	Mapped f;
	if (!f.open_file(path, error)) return false;
	// binary when the size matches the record count (ASCII files start with "solid", but so do some binary ones)
	if (f.size >= 84) {
		uint32_t n = 0;
		memcpy(&n, f.base + 80, 4);
		if (f.size == 84 + (size_t)n * 50) {
			m.vertices.resize((size_t)n * 9);
			m.faces.resize((size_t)n * 3);
			for (size_t i = 0; i < n; i++) {
				memcpy(m.vertices.data() + 9 * i, f.base + 84 + 50 * i + 12, 36);
				for (int k = 0; k < 3; k++) m.faces[3 * i + k] = (int32_t)(3 * i + k);
			}
			return true;
		}
	}
	const char* p = f.base; const char* end = f.base + f.size;
	const char *t, *te;
	while (next_token(p, end, t, te)) {
		if (te - t == 6 && memcmp(t, "vertex", 6) == 0)
			for (int k = 0; k < 3; k++) {
				if (!next_token(p, end, t, te)) { error = "STL file ends inside a vertex"; return false; }
				const char* q = t; m.vertices.push_back(parse_float(q, te));
			}
	}
	const size_t nv = m.vertices.size() / 3;
	if (nv == 0 || nv % 3 != 0) { error = "not an STL file (no facets found)"; return false; }
	m.faces.resize(nv);
	for (size_t i = 0; i < nv; i++) m.faces[i] = (int32_t)i;
	return true;
}

bool load_3ds(const std::string& path, Mesh& m, std::string& error) {
	Mapped f;
	if (!f.open_file(path, error)) return false;
	const unsigned char* b = reinterpret_cast<const unsigned char*>(f.base);
	auto u16 = [&](size_t at) { return (unsigned)(b[at] | (b[at + 1] << 8)); };
	auto u32 = [&](size_t at) { return (size_t)b[at] | ((size_t)b[at + 1] << 8) | ((size_t)b[at + 2] << 16) | ((size_t)b[at + 3] << 24); };
	if (f.size < 6 || u16(0) != 0x4D4D) { error = "not a 3DS file"; return false; }
	size_t vertex_base = 0;
	// iterative walk: container chunks are entered, everything else is skipped by its length
	size_t at = 0;
	while (at + 6 <= f.size) {
		const unsigned id = u16(at);
		const size_t len = u32(at + 2);
		if (len < 6 || at + len > f.size) { error = "3DS chunk runs past the end of the file"; return false; }
		if (id == 0x4D4D || id == 0x3D3D || id == 0x4100) { at += 6; continue; }                    // main, editor, triangle mesh: descend
		if (id == 0x4000) { size_t q = at + 6; while (q < at + len && b[q]) q++; at = q + 1; continue; }   // object: skip its name, descend
		if (id == 0x4110) {
			const size_t n = u16(at + 6);
			if (8 + n * 12 > len) { error = "3DS vertex list is truncated"; return false; }
			vertex_base = m.vertices.size() / 3;
			const size_t old = m.vertices.size();
			m.vertices.resize(old + n * 3);
			memcpy(m.vertices.data() + old, b + at + 8, n * 12);
		} else if (id == 0x4120) {
			const size_t n = u16(at + 6);
			if (8 + n * 8 > len) { error = "3DS face list is truncated"; return false; }
			for (size_t i = 0; i < n; i++)
				for (int k = 0; k < 3; k++) m.faces.push_back((int32_t)(vertex_base + u16(at + 8 + 8 * i + 2 * k)));
		}
		at += len;
	}
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
	else if (ext == "off") ok = load_off(path, m, error);
	else if (ext == "stl") ok = load_stl(path, m, error);
	else if (ext == "3ds") ok = load_3ds(path, m, error);
	else { error = "unsupported mesh format ." + ext + " (this build reads .obj, .ply, .off, .stl and .3ds; trimesh2 is not linked)"; return false; }
	if (!ok) return false;
	if (m.vertices.empty()) { error = "mesh has no vertices"; return false; }
	const int32_t nv = (int32_t)m.n_vertices();
	for (int32_t idx : m.faces)
		if (idx < 0 || idx >= nv) { error = "face index out of range"; return false; }
	finish_bbox(m);
	return true;
}

}  // namespace voxcli
