// VTK XML snapshots without the VTK library: .vts (vtkStructuredGrid) for cubic bodies and .vtu
// (vtkUnstructuredGrid) for simplex bodies, with the fields of the reference's VtkSnapshotter
// (util/snapshot/VtkSnapshotter.hpp:20-61): every 3-vector of the model ("Velocity"), the task's
// vtkSnapshotter.quantitiesToSnap as Float32 scalars and "material_index"; points in VTK order (x fastest,
// grid/cubic/CubicGrid.hpp:34,45-46).  Data go into one raw appended block (UInt32 byte counts, no compression):
// the reference's files come out of vtkXMLWriter (base64 + zlib by default), so the two are the same data model
// in two encodings of the same format, both read by ParaView/VTK.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

#include "gcmb_host.hpp"

namespace gcmb {
namespace vtk {

namespace {

struct Block {
	std::string xml;                  // the <DataArray .../> element without the offset
	std::vector<unsigned char> bytes;
};

template<typename T>
Block dataArray(const std::string& type, const std::string& name, int components, const std::vector<T>& data) {
	Block b;
	b.xml = "<DataArray type=\"" + type + "\" Name=\"" + name + "\" NumberOfComponents=\"" + std::to_string(components) +
	        "\" format=\"appended\"";
	b.bytes.resize(data.size() * sizeof(T));
	if (!data.empty()) { std::memcpy(b.bytes.data(), data.data(), b.bytes.size()); }
	return b;
}

struct File {
	FILE* f;
	std::vector<Block> blocks;
	size_t offset = 0;
	explicit File(const std::string& path) : f(std::fopen(path.c_str(), "wb")) {
		if (!f) { throw Exception(GCMB_E_INVALID_OP, "cannot open " + path); }
	}
	~File() { if (f) { std::fclose(f); } }
	void text(const std::string& s) { std::fputs(s.c_str(), f); }
	void array(const Block& b) {
		text("        " + b.xml + " offset=\"" + std::to_string(offset) + "\"/>\n");
		offset += sizeof(uint32_t) + b.bytes.size();
		blocks.push_back(b);
	}
	void appended() {
		text("  <AppendedData encoding=\"raw\">\n   _");
		for (const Block& b : blocks) {
			const uint32_t n = (uint32_t) b.bytes.size();
			std::fwrite(&n, sizeof n, 1, f);
			if (n) { std::fwrite(b.bytes.data(), 1, n, f); }
		}
		text("\n  </AppendedData>\n");
	}
};

void pointData(File& out, const std::vector<Field>& fields) {
	out.text("      <PointData>\n");
	for (const Field& fd : fields) { out.array(dataArray("Float32", fd.name, fd.components, fd.data)); }
	out.text("      </PointData>\n");
}

}  // namespace

std::string snapshotFileName(const std::string& outputDirectory, const std::string& folder, size_t meshId, int rank, int step,
		const std::string& extension) {
	// Snapshotter::makeFileNameForSnapshot (util/snapshot/Snapshotter.hpp:53-68)
	std::string dir = "snapshots";
	mkdir(dir.c_str(), 0777);
	if (!outputDirectory.empty()) { dir += "/" + outputDirectory; mkdir(dir.c_str(), 0777); }
	dir += "/" + folder;
	mkdir(dir.c_str(), 0777);
	char name[128];
	std::snprintf(name, sizeof name, "mesh%zucore%02dsnap%04d.%s", meshId, rank, step, extension.c_str());
	return dir + "/" + name;
}

void writeStructuredGrid(const std::string& path, const int (&n)[3], const std::vector<float>& points,
		const std::vector<Field>& fields) {
	File out(path);
	const std::string extent = "0 " + std::to_string(n[0] - 1) + " 0 " + std::to_string(n[1] - 1) + " 0 " + std::to_string(n[2] - 1);
	out.text("<?xml version=\"1.0\"?>\n<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n");
	out.text("  <StructuredGrid WholeExtent=\"" + extent + "\">\n    <Piece Extent=\"" + extent + "\">\n");
	pointData(out, fields);
	out.text("      <CellData>\n      </CellData>\n      <Points>\n");
	out.array(dataArray("Float32", "Points", 3, points));
	out.text("      </Points>\n    </Piece>\n  </StructuredGrid>\n");
	out.appended();
	out.text("</VTKFile>\n");
}

void writeUnstructuredGrid(const std::string& path, const std::vector<float>& points, const std::vector<int32_t>& tetrahedra,
		const std::vector<Field>& fields) {
	File out(path);
	const size_t nPoints = points.size() / 3, nCells = tetrahedra.size() / 4;
	std::vector<int32_t> offsets(nCells);
	std::vector<uint8_t> types(nCells, 10);   // VTK_TETRA
	for (size_t c = 0; c < nCells; c++) { offsets[c] = (int32_t) (4 * (c + 1)); }
	out.text("<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n");
	out.text("  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"" + std::to_string(nPoints) + "\" NumberOfCells=\"" + std::to_string(nCells) + "\">\n");
	pointData(out, fields);
	out.text("      <CellData>\n      </CellData>\n      <Points>\n");
	out.array(dataArray("Float32", "Points", 3, points));
	out.text("      </Points>\n      <Cells>\n");
	out.array(dataArray("Int32", "connectivity", 1, tetrahedra));
	out.array(dataArray("Int32", "offsets", 1, offsets));
	out.array(dataArray("UInt8", "types", 1, types));
	out.text("      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n");
	out.appended();
	out.text("</VTKFile>\n");
}

/// the fields VtkSnapshotter::snapshotImpl adds, from PDE vectors in `order` (order[i] = index of the i-th VTK
/// point in the pde array)
std::vector<Field> snapshotFields(Models::T model, int D, int M, const std::vector<real>& pde, const std::vector<size_t>& order,
		const std::vector<PhysicalQuantities::T>& quantities, const std::vector<float>& materialIndex) {
	static const char* NAMES[] = {"Velocity", "Force", "Vx", "Vy", "Vz", "Sxx", "Sxy", "Sxz", "Syy", "Syz", "Szz", "rho",
	                              "pressure", "damage_measure"};   // util/Enum.cpp:6-24
	std::vector<Field> fields;
	const size_t n = order.size();
	Field velocity;
	velocity.name = NAMES[0];
	velocity.components = 3;
	velocity.data.assign(3 * n, 0.0f);
	for (size_t i = 0; i < n; i++) for (int d = 0; d < D; d++) { velocity.data[3 * i + (size_t) d] = (float) pde[order[i] * (size_t) M + (size_t) d]; }
	fields.push_back(velocity);
	for (const PhysicalQuantities::T q : quantities) {
		Field fd;
		fd.name = NAMES[(int) q];
		fd.components = 1;
		fd.data.resize(n);
		const int code = quantityCode(model, D, q);
		for (size_t i = 0; i < n; i++) {
			const real* u = &pde[order[i] * (size_t) M];
			real value;
			if (code >= 0) { value = u[code]; }
			else {
				// PdeVariables::GetPressure: -trace(sigma) / D (rheology/variables/VelocitySigmaVariables.hpp)
				real trace = 0;
				for (int d = 0; d < D; d++) { trace += u[sigmaComponent(D, d, d)]; }
				value = -trace / D;
			}
			fd.data[i] = (float) value;
		}
		fields.push_back(fd);
	}
	Field material;
	material.name = "material_index";
	material.components = 1;
	material.data = materialIndex;
	fields.push_back(material);
	return fields;
}

}  // namespace vtk
}  // namespace gcmb
