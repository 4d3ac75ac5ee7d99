// mesh_io.cpp — mesh loading for the CLI (stands in for trimesh2's TriMesh::read + need_faces + need_bbox,
// main.cpp:174-179).  Floats are parsed with strtof, i.e. correctly rounded binary32 like trimesh2's sscanf("%f").
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

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
	if (ext == "obj") ok = load_obj(path, m, error);
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
