// Registry of the stage kernels: the sparsity patterns (patterns.inc) and, per kernel set, the launchers the
// translation units of stage_inst.cu register at load time.
#include <cstring>

#include "internal.cuh"

namespace gcmb {

#define GCMB_L(...) {__VA_ARGS__}
#define GCMB_PATTERN(GROUP, NAME, MM, AXIS, SGN, UM, U1M, BASE, UNEG, U1NEG) \
	{#NAME, GROUP, MM, AXIS, SGN, UM, U1M, BASE, UNEG, U1NEG},
static const PatternInfo g_patterns[] = {
#include "patterns.inc"
};
#undef GCMB_PATTERN

constexpr int N_PATTERNS = (int) (sizeof(g_patterns) / sizeof(g_patterns[0]));

namespace {
struct Registry {
	StageLauncher sparse[N_SETS][N_PATTERNS][N_VARIANTS];
	StageLauncher dense[N_SETS][MAXM + 1];
	StageLauncher dense_k0[N_SETS][MAXM + 1][3][2];
	Registry() { std::memset(this, 0, sizeof *this); }
};
// constructed on first use: the registering translation units' static initialisers may run before this one's
Registry& registry() {
	static Registry r;
	return r;
}
}  // namespace

int pattern_count() { return N_PATTERNS; }
const PatternInfo& pattern(int i) { return g_patterns[i]; }

void register_sparse_launcher(int set, const char* pattern_name, int variant, StageLauncher f) {
	for (int p = 0; p < N_PATTERNS; p++) {
		if (std::strcmp(g_patterns[p].name, pattern_name) == 0) { registry().sparse[set][p][variant] = f; }
	}
}
void register_dense_launcher(int set, int M, StageLauncher f) { registry().dense[set][M] = f; }
void register_dense_k0_launcher(int set, int M, int bs, bool k0rt, StageLauncher f) { registry().dense_k0[set][M][bs][k0rt ? 1 : 0] = f; }

StageLauncher sparse_launcher(int set, int p, int variant) {
	if (set < 0 || set >= N_SETS || p < 0 || p >= N_PATTERNS || variant < 0 || variant >= N_VARIANTS) { return nullptr; }
	return registry().sparse[set][p][variant];
}
StageLauncher dense_launcher(int set, int M) {
	if (set < 0 || set >= N_SETS || M < 1 || M > MAXM) { return nullptr; }
	return registry().dense[set][M];
}
StageLauncher dense_k0_launcher(int set, int M, int bs, bool k0rt) {
	if (set < 0 || set >= N_SETS || M < 1 || M > MAXM || bs < 1 || bs > 2) { return nullptr; }
	return registry().dense_k0[set][M][bs][k0rt ? 1 : 0];
}

}  // namespace gcmb
