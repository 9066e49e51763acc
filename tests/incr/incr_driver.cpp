// incr_driver.cpp - TEST INFRASTRUCTURE: an embedding program that uses SeqFROST's incremental library API
// (Solver() + iadd / itoClause / ifreeze / isolve, reference solve.hpp:797-819, solveinc.cpp:80-116) the way a client of
// libseqfrost.a does.  Built twice from this one source:
//   (default)      against the UNMODIFIED reference: Solver + isolve()
//   -DUSE_SIGMA    against the binding: SigmaSolver + sigmaISolve()  (seqfrost_b200/host/sigma_solver.hpp)
// Both print the same digest lines per isolve() call; tests/test_integration.py compares them.
//
// usage: incr_driver <file.cnf> [reference options] [--freeze=v,v,...] --assume=l,l,... [--assume=...]...
//        incr_driver <file.cnf> [reference options] --solve [--bound]
// --solve: the non-incremental client (Solver(path) + solve()).  --bound switches on opts.boundsearch_en, the flag behind
// --conflictout / --decisionout (solve.hpp:64-66): the reference's own -boundsearch option never reaches it (options.cpp
// registers the option but OPTION::init does not copy it), so a client that wants a deterministic search bound sets the
// field itself - on both builds alike.
#ifdef USE_SIGMA
#include "../../seqfrost_b200/host/sigma_solver.hpp"
#else
#include "control.hpp"
#include "solve.hpp"
#include "can.hpp"
#include "version.hpp"
#endif
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace SeqFROST;

bool quiet_en = false;
bool competition_en = false;
int verbose = -1;

namespace {
#ifdef USE_SIGMA
typedef SigmaSolver Base;
#else
typedef Solver Base;
#endif

class Client : public Base {
public:
	Client() : Base() {}
	explicit Client(const std::string& path) : Base(path) {}
	void boundSearch() { opts.boundsearch_en = true; }
	void plainSolve() {
#ifdef USE_SIGMA
		sigmaSolve();
#else
		solve();
#endif
	}
	bool load(const char* path) {                       // clause by clause through iadd / itoClause (addinc.cpp:33-150)
		FILE* f = fopen(path, "r");
		if (!f) return false;
		Lits_t in_c, org;
		char tok[64];
		int c;
		while ((c = fgetc(f)) != EOF) {
			if (c == 'c' || c == 'p') { while ((c = fgetc(f)) != EOF && c != '\n'); continue; }
			if (c == ' ' || c == '\n' || c == '\r' || c == '\t') continue;
			int n = 0;
			do { if (n < 62) tok[n++] = (char)c; c = fgetc(f); } while (c != EOF && c != ' ' && c != '\n' && c != '\r' && c != '\t');
			tok[n] = 0;
			const long x = atol(tok);
			if (x == 0) { if (!itoClause(in_c, org)) { fclose(f); return false; } continue; }
			const uint32 v = (uint32)(x < 0 ? -x : x);
			while (v > inf.maxVar) iadd();
			org.push(V2DEC(v, LIT_ST(x < 0)));
		}
		fclose(f);
		return true;
	}
	void call(Lits_t& assumptions) {
#ifdef USE_SIGMA
		sigmaISolve(assumptions);
#else
		isolve(assumptions);
#endif
	}
	void digest(const int round) {
		printf("incr round=%d verdict=%s conflicts=%lld decisions=%lld propagations=%lld melted=%u simplify_calls=%d\n", round,
			SAT ? "SAT" : (UNSAT ? "UNSAT" : "UNKNOWN"), (long long)stats.conflicts, (long long)stats.decisions.single,
			(long long)stats.searchprops, inf.maxMelted, (int)stats.simplify.calls);
		fflush(stdout);
	}
};

std::vector<long> csv(const std::string& s) {
	std::vector<long> out;
	size_t p = 0;
	while (p < s.size()) {
		size_t q = s.find(',', p);
		if (q == std::string::npos) q = s.size();
		if (q > p) out.push_back(atol(s.substr(p, q - p).c_str()));
		p = q + 1;
	}
	return out;
}
} // namespace

int main(int argc, char** argv)
{
	BOOL_OPT opt_competition_en("competition", "engage SAT competition mode", false);
	BOOL_OPT opt_quiet_en("quiet", "enable quiet mode, same as verbose=0", false);
	INT_OPT opt_verbose("verbose", "set the verbosity", 1, INT32R(0, 4));
	INT_OPT opt_timeout("timeout", "set out-of-time limit in seconds", 0, INT32R(0, INT32_MAX));
	INT_OPT opt_memoryout("memoryout", "set out-of-memory limit in gigabytes", 0, INT32R(0, 256));
	if (argc < 2) { fprintf(stderr, "usage: incr_driver <cnf> [options] [--freeze=..] --assume=.. ...\n"); return 1; }
	std::vector<std::vector<long>> rounds;
	std::vector<long> freeze;
	bool plain = false, bound = false;
	std::vector<char*> pass;
	for (int i = 0; i < argc; ++i) {
		std::string a = argv[i];
		if (a.rfind("--assume=", 0) == 0) rounds.push_back(csv(a.substr(9)));
		else if (a.rfind("--freeze=", 0) == 0) freeze = csv(a.substr(9));
		else if (a == "--solve") plain = true;
		else if (a == "--bound") bound = true;
		else pass.push_back(argv[i]);
	}
	int pargc = (int)pass.size();
	parseArguments(pargc, pass.data());
	competition_en = opt_competition_en;
	quiet_en = opt_quiet_en, verbose = opt_verbose;
	if (quiet_en || competition_en) verbose = 0;
	else if (!verbose || competition_en) quiet_en = true;
	if (plain) {
		Client* s = new Client(std::string(argv[1]));
		solver = s;
		if (bound) s->boundSearch();
		s->plainSolve();
		s->digest(0);
		_exit(0);
	}
	Client* s = new Client();
	solver = s;
	if (!s->load(argv[1])) { printf("incr load: UNSAT\n"); fflush(stdout); _exit(0); }
	for (long v : freeze) s->ifreeze((uint32)v);
	int round = 0;
	for (const std::vector<long>& r : rounds) {
		Lits_t assumptions;
		for (long x : r) assumptions.push(V2DEC((uint32)(x < 0 ? -x : x), LIT_ST(x < 0)));
		s->call(assumptions);
		s->digest(round++);
	}
	fflush(stdout);
	_exit(0);
}
