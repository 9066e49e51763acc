// ref_harness.cpp - TEST INFRASTRUCTURE (never linked into the product).
//
// Drives the UNMODIFIED reference (libseqfrost_ref.a, compiled by oracle/Makefile from
// /root/reference/src) through one SIGmA pass, stage by stage, exactly as
// Solver::simplifying() does (reference simplify.cpp:155-233), and writes
//   --sgin=<file>   the drop-in boundary INPUT  (state right after awaken(), simplify.cpp:171)
//   --cmout=<file>  the solver's clause database right AFTER newBeginning() without variable mapping (simplify.cpp:249-268):
//                   CMM arena, orgs, learnts, stats.literals, sp->state[v].subsume before / after
//   --cmin=<file>   the solver's clause database right BEFORE awaken() (input of extract(), simplify.cpp:118-153):
//                   CMM arena, orgs, learnts, deleted stencil
//   --sgout=<file>  the drop-in boundary OUTPUT (state right after shrinkSimp(), simplify.cpp:210)
//   --trace=<file>  per-phase elected lists / counters (debug aid for stage-level parity)
// in the sectioned binary format read by seqfrost_b200/boundary.py.  It is also the
// "reference" CPU baseline: it times the replaced region (simplify.cpp:180-206) on one core.
//
//   --incr=seed:permille:nassume   incremental solver (Solver() + iadd/itoClause): isolve()'s prologue, a seeded subset of the variables
//                   frozen through ifreeze() plus `nassume` assumptions; the .in then carries the `ifrozen` mask
//   --cmmap=<file>  (instead of --cmout) the same after map(true): the variable mapping, the mapped clause database, the watch table
//   --parseout=<file>  only the reference's DIMACS parser (dimacs.cpp:25-177): CMM arena, orgs, parse-time units (trail), value[],
//                   state[], header and formula counters, dumped right after parser() returns
//   --full-timing   also time awaken() (extract), newBeginning()/map(true) and rebuildWT(): the whole simplifying() call
// usage: ref_sigma <file.cnf> [reference CLI options] [--sgin=..] [--sgout=..] [--trace=..] [--incr=..]
//
// All members touched are public/protected in the reference's Solver (solve.hpp:75-844);
// the reference sources are not modified.
#include "control.hpp"
#include "solve.hpp"
#include "version.hpp"
#include "can.hpp"
#include <chrono>
#include <cstdio>
#include <string>
#include <vector>

using namespace SeqFROST;

bool quiet_en = false;
bool competition_en = false;
int verbose = -1;

namespace {

struct Writer {
	FILE* f = nullptr;
	bool open(const std::string& p) { f = fopen(p.c_str(), "wb"); if (!f) return false; fwrite("SGMA0001", 1, 8, f); return true; }
	void section(const char* name, const void* data, uint64_t nbytes) {
		char nm[8] = { 0 };
		strncpy(nm, name, 8);
		fwrite(nm, 1, 8, f);
		fwrite(&nbytes, 8, 1, f);
		if (nbytes) fwrite(data, 1, nbytes, f);
	}
	void close() { if (f) fclose(f), f = nullptr; }
};

std::string g_sgout;
bool g_out_written = false;

// killSolver() (control.cpp:221) ends the process through exit(); when that happens inside the
// pass (UNSAT found by prop/SUB/VE) record the verdict so the parity tests can compare it.
void write_unsat_at_exit() {
	if (g_out_written || g_sgout.empty()) return;
	Writer w;
	if (!w.open(g_sgout)) return;
	uint64_t meta[16] = { 0 };
	meta[0] = 1; // status: 1 = UNSAT / process ended inside the pass
	w.section("meta", meta, sizeof(meta));
	w.close();
}

class Probe : public Solver {
public:
	Probe(const std::string& path) : Solver(path) {}
	Probe() : Solver() {}                 // the library's incremental constructor (solveinc.cpp:26-49)

	// incremental front end: feeds a DIMACS file through the public incremental API, variable by variable and clause by
	// clause (iadd / itoClause, addinc.cpp:33-150), as an embedding program would.  (The reference's own -parseincr path
	// is not usable: dimacs.cpp:94 calls makeClause() before it looks at the option and stops with "too many variables".)
	bool loadIncremental(const std::string& path) {
		FILE* f = fopen(path.c_str(), "r");
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

	void dumpStore(Writer& w) {
		// SCNF arena verbatim (sclause.hpp:45-49 AoS, simptypes.hpp:48-95) + refs in allocation order
		uint32* base = (uint32*)scnf.clause(0);
		const uint64_t buckets = scnf.STYPE::size(); // bump pointer of the arena (memory.hpp:94)
		const uint32 n = scnf.size();
		w.section("arena", base, buckets * sizeof(uint32));
		w.section("refs", scnf.refs().data(), uint64_t(n) * sizeof(S_REF));
	}

	void dumpCommon(Writer& w, uint64_t status, double passms) {
		uint64_t meta[16] = { 0 };
		meta[0] = status;
		meta[1] = inf.maxVar;
		meta[2] = inf.nClauses;
		meta[3] = inf.nLiterals;
		meta[4] = inf.unassigned;
		meta[5] = inf.maxFrozen;
		meta[6] = inf.maxMelted;
		meta[7] = trail.size();
		meta[8] = sp->propagated;
		meta[9] = stats.simplify.calls;
		meta[10] = (uint64_t)(passms * 1000.0); // microseconds in replaced region
		meta[11] = inf.orgVars;
		meta[12] = (uint64_t)phase;
		w.section("meta", meta, sizeof(meta));
		std::vector<uint8_t> st(inf.maxVar + 1);
		for (uint32 v = 0; v <= inf.maxVar; ++v) st[v] = (uint8_t)sp->state[v].state;
		w.section("state", st.data(), st.size());
		w.section("value", sp->value, inf.nDualVars);
		w.section("trail", trail.data(), uint64_t(trail.size()) * 4);
		w.section("vorg", vorg.data(), uint64_t(vorg.size()) * 4);
		w.section("model", model.resolved.data(), uint64_t(model.resolved.size()) * 4);
		if (incremental && ifrozen.size() >= inf.maxVar + 1)   // the mask iassumed() reads (solve.hpp:816)
			w.section("ifrozen", ifrozen.data(), uint64_t(inf.maxVar) + 1);
		int32_t o[32] = { 0 };
		o[0] = opts.phases; o[1] = opts.phase_lits_min; o[2] = opts.mu_pos; o[3] = opts.mu_neg;
		o[4] = opts.lcve_max_occurs; o[5] = opts.lcve_clause_max; o[6] = opts.lcve_min_vars;
		o[7] = opts.sub_max_occurs; o[8] = opts.sub_clause_max; o[9] = opts.ve_clause_max;
		o[10] = opts.xor_max_arity; o[11] = opts.bce_max_occurs; o[12] = opts.ere_max_occurs;
		o[13] = opts.ere_clause_max; o[14] = opts.collect_freq; o[15] = opts.lbd_tier1;
		o[16] = opts.sub_en; o[17] = opts.ve_en; o[18] = opts.ve_plus_en; o[19] = opts.ve_fun_en;
		o[20] = opts.ve_lbound_en; o[21] = opts.bce_en; o[22] = opts.ere_en; o[23] = opts.ere_extend;
		o[24] = opts.all_en; o[25] = opts.proof_en;
		w.section("opts", o, sizeof(o));
#ifdef STATISTICS
		int64_t s[16] = { 0 };
		s[0] = stats.simplify.bve.pures; s[1] = stats.simplify.bve.inverters; s[2] = stats.simplify.bve.andors;
		s[3] = stats.simplify.bve.ites; s[4] = stats.simplify.bve.xors; s[5] = stats.simplify.bve.funs;
		s[6] = stats.simplify.bve.resolutions; s[7] = stats.simplify.sub.subsumed; s[8] = stats.simplify.sub.strengthened;
		w.section("stats", s, sizeof(s));
#endif
	}

	// optional prologue: the reference's own first simplify() followed by `conflicts` conflicts of its CDCL
	// loop (solve.cpp:161-175 without the inprocessing branches), so that the staged pass below is a LATER
	// call: learnt clauses in the store, stats.simplify.calls > 1 (countOrgs, bumpShrunken paths)
	bool presolve(const int64 conflicts) {
		if (canPreSimplify()) simplify();
		if (!UNSOLVED) return false;
		MDMInit();
		while (UNSOLVED && (int64)stats.conflicts < conflicts) {
			if (BCP()) analyze();
			else if (!inf.unassigned) SET_SAT;
			else if (canReduce()) reduce();
			else if (canRestart()) restart();
			else if (canRephase()) rephase();
			else if (canMMD()) MDM();
			else decide();
		}
		return UNSOLVED;
	}

	// --parseout: the reference's own DIMACS parser (dimacs.cpp:25-177) on `path`, through the library constructor so that the
	// dump is taken right after parser() returns - before the BCP() the file constructor (solve.cpp:52) runs next.
	int parseOnly(const std::string& path, const std::string& outp) {
		formula = FORMULA(path);
		const bool ok = parser();
		Writer w;
		if (!w.open(outp)) return 5;
		uint64_t meta[16] = { 0 };
		meta[0] = ok ? 1 : 0;
		meta[1] = inf.orgVars; meta[2] = inf.orgCls; meta[3] = inf.maxVar;
		meta[4] = formula.units; meta[5] = formula.binaries; meta[6] = formula.ternaries; meta[7] = formula.large;
		meta[8] = (uint64_t)formula.maxClauseSize; meta[9] = stats.literals.original; meta[10] = stats.clauses.original;
		meta[11] = trail.size(); meta[12] = inf.unassigned; meta[13] = inf.maxFrozen;
		w.section("meta", meta, sizeof meta);
		if (ok) {
			w.section("cm", cm.address(0), uint64_t(cm.size()) * sizeof(uint32));
			w.section("orgs", orgs.data(), uint64_t(orgs.size()) * sizeof(C_REF));
			w.section("trail", trail.data(), uint64_t(trail.size()) * sizeof(uint32));
			w.section("value", sp->value, uint64_t(inf.nDualVars));
			std::vector<uint8_t> st(inf.maxVar + 1);
			for (uint32 v = 0; v <= inf.maxVar; ++v) st[v] = sp->state[v].state;
			w.section("state", st.data(), st.size());
		}
		w.close();
		return 0;
	}

	void dumpClauseDb(Writer& w) {
		// CMM arena verbatim (solvetypes.hpp:57, clause.hpp:47-59) + the two ref lists + cm.stencil() up to the last ref
		const uint64_t buckets = cm.size();
		w.section("cm", cm.address(0), buckets * sizeof(uint32));
		w.section("orgs", orgs.data(), uint64_t(orgs.size()) * sizeof(C_REF));
		w.section("learnts", learnts.data(), uint64_t(learnts.size()) * sizeof(C_REF));
		uint64_t top = 0;
		for (uint32 i = 0; i < orgs.size(); ++i) if (orgs[i] + 1 > top) top = orgs[i] + 1;
		for (uint32 i = 0; i < learnts.size(); ++i) if (learnts[i] + 1 > top) top = learnts[i] + 1;
		w.section("stencil", cm.stencil(), top);
	}

	// incremental mode (--incr=seed:permille:nassume; the formula came in through iadd()/itoClause()): the prologue of isolve() (solveinc.cpp:82-93) with a seeded subset of the
	// variables frozen through ifreeze() (assume.cpp:45) and `nassume` assumptions through iassume() (assume.cpp:68), so
	// that LCVE skips them (iassumed, lcve.cpp:67, solve.hpp:816).  Returns false when BCP finds a root conflict.
	bool incrPrologue(const uint64_t seed, const uint32 permille, const uint32 nassume) {
		iallocSpace();
		iunassume();
		if (BCP()) return false;
		initLimits();
		uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
		auto next = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
		Lits_t assumed;
		for (uint32 v = 1; v <= inf.maxVar; ++v) {
			const uint64_t r = next();
			if (sp->state[v].state || !UNASSIGNED(sp->value[V2L(v)])) continue;
			if (r % 1000 < permille) ifreeze(v);
			else if ((uint32)assumed.size() < nassume && (r >> 20) % 97 == 0) assumed.push(V2DEC(v, LIT_ST((r >> 40) & 1)));
		}
		iassume(assumed);
		return true;
	}

	int run(const std::string& sgin, const std::string& sgout, const std::string& tracep, const int64 presolve_conflicts, const std::string& cmin = std::string(),
		const std::string& cmout = std::string(), const std::string& incr = std::string(), const bool fulltiming = false,
		const std::string& cmmap = std::string()) {
		using clk = std::chrono::steady_clock;
		timer.start();
		if (!incr.empty()) {
			unsigned long long seed = 1; unsigned permille = 100, nassume = 0;
			sscanf(incr.c_str(), "%llu:%u:%u", &seed, &permille, &nassume);
			if (!incrPrologue(seed, permille, nassume)) { fprintf(stderr, "ref_sigma: incremental prologue failed\n"); return 7; }
		}
		else
		initLimits();
		if (presolve_conflicts > 0 && !presolve(presolve_conflicts)) { fprintf(stderr, "ref_sigma: solved during the prologue\n"); return 6; }
		if (!opts.preprocess_en) { fprintf(stderr, "ref_sigma: preprocess disabled\n"); return 2; }
		if (!opts.phases && !(opts.all_en || opts.ere_en)) { fprintf(stderr, "ref_sigma: no phases\n"); return 2; }
		stats.simplify.calls++;            // simplify.cpp:108
		rootify();                         // simplify.cpp:161
		shrinkTop(false);                  // simplify.cpp:162
		if (orgs.empty()) { fprintf(stderr, "ref_sigma: formula empty after root shrinking\n"); return 3; }
		if (!cmin.empty()) {
			Writer w;
			if (!w.open(cmin)) return 5;
			dumpClauseDb(w);
			w.close();
		}
		auto ta0 = clk::now();
		awaken();                          // simplify.cpp:171
		auto ta1 = clk::now();
		if (simpstate == AWAKEN_FAIL) return 4;
		const uint64_t literals_in = inf.nLiterals;   // after extract(): the "literals/s" numerator (simplify.cpp:146-150)
		if (!sgin.empty()) {
			Writer w;
			if (!w.open(sgin)) return 5;
			dumpCommon(w, 0, 0.0);
			dumpStore(w);
			w.close();
		}
		Writer tw;
		const bool tracing = !tracep.empty() && tw.open(tracep);
		// ---- replaced region: simplify.cpp:178-206 ----
		auto t0 = clk::now();
		int64 litsbefore = inf.nLiterals, diff = INT64_MAX;
		while (litsbefore) {
			resizeCNF();
			createOT();
			if (!prop()) killSolver();
			if (!LCVE()) break;
			sortOT();
			if (tracing) {
				char nm[16];
				snprintf(nm, sizeof nm, "elect%d", phase);
				tw.section(nm, elected.data(), uint64_t(elected.size()) * 4);
			}
			if ((phase == opts.phases) || (diff <= opts.phase_lits_min && phase > 2)) { ERE(); break; }
			SUB(), VE(), BCE();
			countAll(), filterElected();
			inf.nClauses = inf.nClausesAfter, inf.nLiterals = inf.nLiteralsAfter;
			diff = litsbefore - inf.nLiterals, litsbefore = inf.nLiterals;
			if (tracing) {
				char nm[16];
				snprintf(nm, sizeof nm, "count%d", phase);
				uint64_t c[4] = { inf.nClauses, inf.nLiterals, trail.size(), model.resolved.size() };
				tw.section(nm, c, sizeof c);
			}
			phase++;
			multiplier++;
			multiplier += phase == opts.phases;
		}
		if (!prop()) killSolver();
		auto t1 = clk::now();
		// ---- write back prologue: simplify.cpp:207-210 ----
		occurs.clear(true), ot.clear(true);
		clearMapFrozen();
		countFinal();
		shrinkSimp();
		auto t2 = clk::now();
		const double passms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		const double gcms = std::chrono::duration<double, std::milli>(t2 - t1).count();
		if (tracing) tw.close();
		if (!sgout.empty()) {
			Writer w;
			if (!w.open(sgout)) return 5;
			dumpCommon(w, 0, passms);
			dumpStore(w);
			w.close();
		}
		g_out_written = true;
		if (!cmout.empty() && inf.unassigned && inf.nClauses) {
			// simplify.cpp:215-217 with canMap() false: newBeginning() builds cm / orgs / learnts from the compacted SCNF
			std::vector<uint8_t> s0(inf.maxVar + 1), s1(inf.maxVar + 1);
			for (uint32 v = 0; v <= inf.maxVar; ++v) s0[v] = sp->state[v].subsume;
			newBeginning();
			for (uint32 v = 0; v <= inf.maxVar; ++v) s1[v] = sp->state[v].subsume;
			Writer w;
			if (!w.open(cmout)) return 5;
			w.section("cm", cm.address(0), uint64_t(cm.size()) * sizeof(uint32));
			w.section("orgs", orgs.data(), uint64_t(orgs.size()) * sizeof(C_REF));
			w.section("learnts", learnts.data(), uint64_t(learnts.size()) * sizeof(C_REF));
			uint64_t m[4] = { stats.literals.original, stats.literals.learnt, stats.clauses.original, stats.clauses.learnt };
			w.section("cmmeta", m, sizeof m);
			w.section("subsume0", s0.data(), s0.size());
			w.section("subsume1", s1.data(), s1.size());
			// rebuildWT(opts.simplify_priorbins) (simplify.cpp:218, watch.cpp:135-157): the watch table as one offsets array +
			// the WATCH records {C_REF ref; uint32 imp; int size} of every list in order, and the arena again (sortClause may
			// have reordered literals)
			rebuildWT(opts.simplify_priorbins);
			std::vector<uint32_t> woff(inf.nDualVars + 1, 0);
			std::vector<WATCH> wrec;
			for (uint32 lit = 0; lit < inf.nDualVars; ++lit) {
				woff[lit] = (uint32_t)wrec.size();
				WL& ws = wt[lit];
				for (uint32 k = 0; k < ws.size(); ++k) wrec.push_back(ws[k]);
			}
			woff[inf.nDualVars] = (uint32_t)wrec.size();
			w.section("wtoff", woff.data(), woff.size() * sizeof(uint32_t));
			w.section("wtrec", wrec.data(), wrec.size() * sizeof(WATCH));
			w.section("cm2", cm.address(0), uint64_t(cm.size()) * sizeof(uint32));
			uint64_t code = (uint64_t)opts.simplify_priorbins;
			w.section("wtcode", &code, sizeof code);
			w.close();
		}
		if (!cmmap.empty() && inf.unassigned && inf.nClauses) {
			// simplify.cpp:215 with canMap() taken as true: map(true) (map.cpp:24-121) - variable mapping, newBeginning() with
			// mapLit() on every literal - then rebuildWT() on the new numbering.  The mapping itself is dumped first, computed
			// exactly as vmap.initiate() computes it (map.hpp:78-95); map(true) destroys its own copy.
			std::vector<uint32_t> vm(inf.maxVar + 1, 0);
			uint32_t newVars = 0, firstDL0 = 0;
			for (uint32 v = 1; v <= inf.maxVar; ++v) {
				if (!sp->state[v].state) vm[v] = ++newVars;
				else if (!firstDL0 && !UNASSIGNED(sp->value[V2L(v)])) { firstDL0 = v; vm[v] = ++newVars; }
			}
			std::vector<uint8_t> s0(inf.maxVar + 1);
			for (uint32 v = 0; v <= inf.maxVar; ++v) s0[v] = sp->state[v].subsume;
			map(true);
			Writer w;
			if (!w.open(cmmap)) return 5;
			w.section("vmap", vm.data(), vm.size() * sizeof(uint32_t));
			uint64_t nv[2] = { newVars, inf.maxVar };
			w.section("vmapn", nv, sizeof nv);
			w.section("cm", cm.address(0), uint64_t(cm.size()) * sizeof(uint32));
			w.section("orgs", orgs.data(), uint64_t(orgs.size()) * sizeof(C_REF));
			w.section("learnts", learnts.data(), uint64_t(learnts.size()) * sizeof(C_REF));
			uint64_t m[4] = { stats.literals.original, stats.literals.learnt, stats.clauses.original, stats.clauses.learnt };
			w.section("cmmeta", m, sizeof m);
			w.section("subsume0", s0.data(), s0.size());
			w.section("subsume1", s0.data(), s0.size());
			rebuildWT(opts.simplify_priorbins);
			std::vector<uint32_t> woff(inf.nDualVars + 1, 0);
			std::vector<WATCH> wrec;
			for (uint32 lit = 0; lit < inf.nDualVars; ++lit) {
				woff[lit] = (uint32_t)wrec.size();
				WL& ws = wt[lit];
				for (uint32 k = 0; k < ws.size(); ++k) wrec.push_back(ws[k]);
			}
			woff[inf.nDualVars] = (uint32_t)wrec.size();
			w.section("wtoff", woff.data(), woff.size() * sizeof(uint32_t));
			w.section("wtrec", wrec.data(), wrec.size() * sizeof(WATCH));
			w.section("cm2", cm.address(0), uint64_t(cm.size()) * sizeof(uint32));
			uint64_t code = (uint64_t)opts.simplify_priorbins;
			w.section("wtcode", &code, sizeof code);
			w.close();
		}
		if (fulltiming && cmout.empty() && inf.unassigned && inf.nClauses) {
			// the rest of simplifying() (simplify.cpp:215-229) for the whole-call figure the CLI prints as "Simplifier time":
			// newBeginning() (or map(true)) and rebuildWT()
			auto t3 = clk::now();
			if (canMap()) map(true);
			else newBeginning();
			auto t4 = clk::now();
			rebuildWT(opts.simplify_priorbins);
			auto t5 = clk::now();
			printf("ref_sigma: awaken_ms=%.3f newbeginning_ms=%.3f rebuildwt_ms=%.3f simplifier_ms=%.3f\n",
				std::chrono::duration<double, std::milli>(ta1 - ta0).count(), std::chrono::duration<double, std::milli>(t4 - t3).count(),
				std::chrono::duration<double, std::milli>(t5 - t4).count(),
				std::chrono::duration<double, std::milli>(ta1 - ta0).count() + passms + gcms + std::chrono::duration<double, std::milli>(t5 - t3).count());
		}
		printf("ref_sigma: pass_ms=%.3f final_compact_ms=%.3f phases=%d vars=%u clauses=%u literals=%u melted=%u units=%u model=%u literals_in=%llu\n",
			passms, gcms, phase, inf.maxVar, inf.nClauses, inf.nLiterals, inf.maxMelted, (unsigned)trail.size(), (unsigned)model.resolved.size(),
			(unsigned long long)literals_in);
		fflush(stdout);
		return 0;
	}
};

} // namespace

int main(int argc, char** argv)
{
	BOOL_OPT opt_competition_en("competition", "engage SAT competition mode", false);
	BOOL_OPT opt_quiet_en("quiet", "enable quiet mode, same as verbose=0", false);
	INT_OPT opt_verbose("verbose", "set the verbosity", 1, INT32R(0, 4));
	INT_OPT opt_timeout("timeout", "set out-of-time limit in seconds", 0, INT32R(0, INT32_MAX));
	INT_OPT opt_memoryout("memoryout", "set out-of-memory limit in gigabytes", 0, INT32R(0, 256));
	if (argc < 2) { fprintf(stderr, "usage: ref_sigma <cnf> [options] [--sgin=f] [--sgout=f] [--trace=f]\n"); return 1; }
	std::string sgin, sgout, tracep, cmin, cmout, incr, parseout, cmmap;
	long long presolve_conflicts = 0;
	bool fulltiming = false;
	std::vector<char*> pass;
	for (int i = 0; i < argc; ++i) {
		std::string a = argv[i];
		if (a.rfind("--sgin=", 0) == 0) sgin = a.substr(7);
		else if (a.rfind("--sgout=", 0) == 0) sgout = a.substr(8);
		else if (a.rfind("--trace=", 0) == 0) tracep = a.substr(8);
		else if (a.rfind("--cmin=", 0) == 0) cmin = a.substr(7);
		else if (a.rfind("--cmout=", 0) == 0) cmout = a.substr(8);
		else if (a.rfind("--cmmap=", 0) == 0) cmmap = a.substr(8);
		else if (a.rfind("--incr=", 0) == 0) incr = a.substr(7);
		else if (a.rfind("--parseout=", 0) == 0) parseout = a.substr(11);
		else if (a == "--full-timing") fulltiming = true;
		else if (a.rfind("--presolve-conflicts=", 0) == 0) presolve_conflicts = atoll(a.substr(21).c_str());
		else pass.push_back(argv[i]);
	}
	g_sgout = sgout;
	atexit(write_unsat_at_exit);
	try {
		int pargc = (int)pass.size();
		parseArguments(pargc, pass.data());
		competition_en = opt_competition_en;
		quiet_en = opt_quiet_en, verbose = opt_verbose;
		if (quiet_en || competition_en) verbose = 0;
		else if (!verbose || competition_en) quiet_en = true;
		std::string formula = argv[1];
		Probe* p = nullptr;
		if (!parseout.empty()) {
			p = new Probe();
			solver = p;
			g_out_written = true;
			const int prc = p->parseOnly(formula, parseout);
			fflush(stdout);
			_exit(prc);
		}
		if (incr.empty()) p = new Probe(formula);
		else {
			p = new Probe();
			solver = p;
			if (!p->loadIncremental(formula)) { fprintf(stderr, "ref_sigma: incremental load failed (empty clause or conflicting units)\n"); g_out_written = true; _exit(8); }
		}
		solver = p;
		int rc = p->run(sgin, sgout, tracep, presolve_conflicts, cmin, cmout, incr, fulltiming, cmmap);
		g_out_written = g_out_written || rc != 0;
		fflush(stdout);
		_exit(rc); // skip the reference's destructors (process-global singletons)
	}
	catch (std::bad_alloc&) { fprintf(stderr, "ref_sigma: bad_alloc\n"); g_out_written = true; return 9; }
	catch (MEMOUTEXCEPTION&) { fprintf(stderr, "ref_sigma: memout\n"); g_out_written = true; return 9; }
}
