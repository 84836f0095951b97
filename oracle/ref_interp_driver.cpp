// TEST INFRASTRUCTURE ONLY: runs the reference's own TriangleInterpolator<real> (header-only,
// src/libgcm/util/math/interpolation/TriangleInterpolator.hpp, unmodified) over queries read from a binary file.
//   gcm_ref_interp <in.bin> <out.bin>
// in.bin : int32 mode, int32 n, then doubles: points [n][3|4][2], values [n][3|4], grads [n][3][2] (modes 1-3), queries [n][2]
// out.bin: doubles out[n], then int32 status[n] (1 = the reference threw)
#include <cstdio>
#include <vector>

#include <libgcm/util/math/interpolation/TriangleInterpolator.hpp>

using namespace gcm;

int main(int argc, char** argv) {
	if (argc < 3) { return 2; }
	FILE* f = fopen(argv[1], "rb");
	if (!f) { return 3; }
	int mode = 0, n = 0;
	if (fread(&mode, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1) { return 4; }
	const int np = mode == 4 ? 4 : 3;
	std::vector<double> points((size_t) n * np * 2), values((size_t) n * np), grads(mode >= 1 && mode <= 3 ? (size_t) n * 6 : 0), queries((size_t) n * 2);
	if (fread(points.data(), 8, points.size(), f) != points.size() || fread(values.data(), 8, values.size(), f) != values.size()) { return 4; }
	if (!grads.empty() && fread(grads.data(), 8, grads.size(), f) != grads.size()) { return 4; }
	if (fread(queries.data(), 8, queries.size(), f) != queries.size()) { return 4; }
	fclose(f);
	typedef TriangleInterpolator<real> TI;
	typedef TI::Gradient G;
	std::vector<double> out((size_t) n, 0.0);
	std::vector<int> status((size_t) n, 0);
	for (int i = 0; i < n; i++) {
		const double* c = points.data() + (size_t) i * np * 2;
		const double* v = values.data() + (size_t) i * np;
		const Real2 q = {queries[2 * (size_t) i], queries[2 * (size_t) i + 1]};
		try {
			if (mode == 4) {
				out[i] = TI::interpolateInOwner({c[0], c[1]}, v[0], {c[2], c[3]}, v[1], {c[4], c[5]}, v[2], {c[6], c[7]}, v[3], q);
				continue;
			}
			const Real2 c0 = {c[0], c[1]}, c1 = {c[2], c[3]}, c2 = {c[4], c[5]};
			if (mode == 0) { out[i] = TI::interpolate(c0, v[0], c1, v[1], c2, v[2], q); continue; }
			const double* g = grads.data() + (size_t) i * 6;
			const G g0({g[0], g[1]}), g1({g[2], g[3]}), g2({g[4], g[5]});
			if (mode == 1) { out[i] = TI::interpolate(c0, v[0], g0, c1, v[1], g1, c2, v[2], g2, q); }
			else if (mode == 2) { out[i] = TI::minMaxInterpolate(c0, v[0], g0, c1, v[1], g1, c2, v[2], g2, q); }
			else { out[i] = TI::hybridInterpolate(c0, v[0], g0, c1, v[1], g1, c2, v[2], g2, q); }
		} catch (Exception&) {
			out[i] = 0;
			status[i] = 1;
		}
	}
	f = fopen(argv[2], "wb");
	if (!f) { return 5; }
	fwrite(out.data(), 8, out.size(), f);
	fwrite(status.data(), 4, status.size(), f);
	fclose(f);
	return 0;
}
