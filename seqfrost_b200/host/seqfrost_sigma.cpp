// seqfrost_sigma.cpp - the reference-side binding of libsigma_b200: SeqFROST with its SIGmA pass
// running on the B200.
//
// Everything except the replaced region is the reference's own code, linked UNMODIFIED from
// libseqfrost(.a): parser, CDCL search, options, model extension/verification, the SCLAUSE hand-off
// (newBeginning -> addClause(SCLAUSE&), sclause.cpp:22-62) and the watch rebuild.  Solver::solve(),
// simplify() and simplifying() are not virtual, so this subclass re-states their few lines of
// top-level control flow (solve.cpp:155-181, simplify.cpp:102-116,155-233) and swaps the phase
// loop simplify.cpp:180-210 for one sigma_run() call.  By default extract() (simplify.cpp:118-133) runs on the
// device as well: awaken() is re-stated without it and the clause database itself is uploaded (sigma_upload_cnf);
// SIGMA_HOST_EXTRACT=1 keeps the reference's awaken() and uploads the SCNF it builds.  Usage is the reference CLI's:
//
//     seqfrost_sigma <formula.cnf> [reference options]
//
// Build: g++ -std=c++17 -O3 -DNDEBUG -DSTATISTICS -I<reference>/src
//     seqfrost_sigma.cpp libseqfrost.a -L. -lsigma_b200
#include "control.hpp"
#include "solve.hpp"
#include "can.hpp"
#include "version.hpp"
#include "../../include/sigma_b200.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace SeqFROST;

bool quiet_en = false;
bool competition_en = false;
int verbose = -1;

namespace {

class SigmaSolver : public Solver {
	sigma_ctx* ctx = nullptr;
	std::vector<uint8_t> st;
	int64 clausesBefore = 0, literalsBefore = 0;      // inf.nClauses / inf.nLiterals after extract()
	bool newBeginningDone = false;                    // cm / orgs / learnts already hold the simplified formula

public:
	explicit SigmaSolver(const std::string& path) : Solver(path) {}

	// awaken() (simplify.cpp:135-153) without extract(): the SCNF and the occurrence table live on the device
	void sigmaAwaken() {
		printStats(1, '-', CGREEN0);
		initSimp();
		wt.clear(true);
		inf.nClauses = inf.nLiterals = 0;
	}

	// simplify.cpp:180-210 on the GPU.  Inputs and outputs are the reference's own objects.
	void sigmaPass(const bool deviceExtract) {
		if (!ctx) {
			const int rc = sigma_create(0, &ctx);
			if (rc != SIGMA_OK) {
				fprintf(stderr, "c sigma_b200: %s (%s)\n", sigma_strerror(rc), ctx ? sigma_last_error(ctx) : "");
				exit(EXIT_FAILURE);                       // no CPU fallback, by design
			}
		}
		sigma_opts o;
		sigma_default_opts(&o);
		o.phases = opts.phases; o.phase_lits_min = opts.phase_lits_min; o.mu_pos = opts.mu_pos; o.mu_neg = opts.mu_neg;
		o.lcve_max_occurs = opts.lcve_max_occurs; o.lcve_clause_max = opts.lcve_clause_max; o.lcve_min_vars = opts.lcve_min_vars;
		o.sub_max_occurs = opts.sub_max_occurs; o.sub_clause_max = opts.sub_clause_max; o.ve_clause_max = opts.ve_clause_max;
		o.xor_max_arity = opts.xor_max_arity; o.bce_max_occurs = opts.bce_max_occurs; o.ere_max_occurs = opts.ere_max_occurs;
		o.ere_clause_max = opts.ere_clause_max; o.collect_freq = opts.collect_freq; o.lbd_tier1 = opts.lbd_tier1;
		o.sub_en = opts.sub_en; o.ve_en = opts.ve_en; o.ve_plus_en = opts.ve_plus_en; o.ve_fun_en = opts.ve_fun_en;
		o.ve_lbound_en = opts.ve_lbound_en; o.bce_en = opts.bce_en; o.ere_en = opts.ere_en; o.ere_extend = opts.ere_extend;
		o.all_en = opts.all_en; o.proof_en = opts.proof_en;

		const uint32 V = inf.maxVar;
		st.resize(V + 1);
		for (uint32 v = 0; v <= V; ++v) st[v] = (uint8_t)sp->state[v].state;
		sigma_formula in = {};
		in.nvars = V;
		if (!deviceExtract) {
			in.n_clauses = scnf.size(); in.n_buckets = scnf.STYPE::size(); in.n_literals = inf.nLiterals;
			in.arena = (uint32_t*)scnf.clause(0); in.refs = (uint64_t*)scnf.refs().data();
		}
		in.state = st.data(); in.value = (int8_t*)sp->value;
		in.trail = trail.data(); in.trail_size = trail.size(); in.propagated = sp->propagated;
		in.vorg = vorg.data(); in.ifrozen = incremental ? (uint8_t*)ifrozen.data() : nullptr;
		in.model = model.resolved.data(); in.model_size = model.resolved.size();
		in.unassigned = inf.unassigned; in.max_frozen = inf.maxFrozen; in.max_melted = inf.maxMelted;
		in.simplify_calls = (uint32_t)stats.simplify.calls;

		int rc;
		if (deviceExtract) {
			// extract(orgs), extract(learnts), cm.destroy() (simplify.cpp:146-149) - the first two on the device
			sigma_cnf db = {};
			db.cm = (const uint32_t*)cm.address(0); db.cm_buckets = cm.size();
			db.orgs = (const uint64_t*)orgs.data(); db.n_orgs = orgs.size();
			db.learnts = (const uint64_t*)learnts.data(); db.n_learnts = learnts.size();
			uint64_t top = 0;                                  // the stencil is valid up to the last allocated ref
			for (uint32 i = 0; i < orgs.size(); ++i) if (orgs[i] >= top) top = orgs[i] + 1;
			for (uint32 i = 0; i < learnts.size(); ++i) if (learnts[i] >= top) top = learnts[i] + 1;
			db.stencil = (const uint8_t*)cm.stencil(); db.stencil_size = top;
			rc = sigma_upload_cnf(ctx, &o, &db, &in);
			if (rc == SIGMA_OK) {
				uint32_t ncls = 0; uint64_t nbuckets = 0, nlits = 0;
				sigma_extracted_sizes(ctx, &ncls, &nbuckets, &nlits);
				inf.nClauses = ncls; inf.nLiterals = (uint32)nlits;
				orgs.clear(true), learnts.clear(true);
				cm.destroy();
				if (verbose >= 2) printf("c  sigma_b200: extracted on the device (%u clauses, %llu literals)\n", ncls, (unsigned long long)nlits);
			}
		}
		else rc = sigma_upload(ctx, &o, &in);
		clausesBefore = inf.nClauses; literalsBefore = inf.nLiterals;
		if (rc == SIGMA_OK) rc = sigma_execute(ctx);
		if (rc == SIGMA_UNSAT) { learnEmpty(); killSolver(); }           // propelim.cpp:60, subsume.cpp:66, bounded.cpp:74
		if (rc != SIGMA_OK) { fprintf(stderr, "c sigma_b200: %s: %s\n", sigma_strerror(rc), sigma_last_error(ctx)); exit(EXIT_FAILURE); }
		sigma_formula sz = {};
		sigma_result_sizes(ctx, &sz);
		inf.nClauses = sz.n_clauses; inf.nLiterals = (uint32)sz.n_literals;
		inf.unassigned = sz.unassigned; inf.maxFrozen = sz.max_frozen; inf.maxMelted = sz.max_melted;
		const uint32 oldtrail = trail.size();
		trail.resize(sz.trail_size);
		model.resolved.resize((uint32)sz.model_size);
		sigma_formula out = {};
		out.state = st.data(); out.value = (int8_t*)sp->value; out.trail = trail.data(); out.model = model.resolved.data();
		out.trail_cap = sz.trail_size; out.model_cap = sz.model_size;
		if (!out.trail) out.trail = (uint32_t*)st.data();
		if (!out.model) out.model = (uint32_t*)st.data();

		// newBeginning() (simplify.cpp:249-268) on the device too when no variable mapping follows (simplify.cpp:215-216):
		// the CLAUSE records, orgs and learnts arrive ready-made
		const char* hn = getenv("SIGMA_HOST_NEWBEGINNING");
		newBeginningDone = false;
		if (!(hn && hn[0] == '1') && inf.unassigned && inf.nClauses && !canMap() && !opts.aggr_cnf_sort) {
			sigma_cnf_out co = {};
			rc = sigma_cnf_out_sizes(ctx, &co);
			if (rc != SIGMA_OK) { fprintf(stderr, "c sigma_b200: %s\n", sigma_last_error(ctx)); exit(EXIT_FAILURE); }
			cm.init(inf.nClauses, inf.nLiterals);
			if (co.last_ref) cm.CTYPE::alloc(co.last_ref);                     // every record in front of the last one, in bulk
			C_REF last = 0;
			cm.alloc(last, int(co.cm_buckets - co.last_ref) - 4);              // the last one through CMM::alloc: it sizes the deleted stencil
			orgs.resize((uint32)co.n_orgs); learnts.resize((uint32)co.n_learnts);
			std::vector<uint8_t> sub(V + 1);
			co.cm = (uint32_t*)cm.address(0); co.cm_cap = co.cm_buckets;
			co.orgs = orgs.size() ? (uint64_t*)orgs.data() : (uint64_t*)st.data(); co.orgs_cap = co.n_orgs;
			co.learnts = learnts.size() ? (uint64_t*)learnts.data() : (uint64_t*)st.data(); co.learnts_cap = co.n_learnts;
			co.subsume = sub.data();
			rc = sigma_download_cnf(ctx, &co, &out);
			if (rc != SIGMA_OK) { fprintf(stderr, "c sigma_b200: download: %s\n", sigma_last_error(ctx)); exit(EXIT_FAILURE); }
			for (uint32 v = 1; v <= V; ++v) if (sub[v]) sp->state[v].subsume = 1;   // mark_ssubsume, sclause.hpp:209-215
			stats.literals.original = co.lits_original; stats.literals.learnt = co.lits_learnt;
			stats.clauses.original = orgs.size(); stats.clauses.learnt = learnts.size();
			scnf.destroy();
			newBeginningDone = true;
			if (verbose >= 2) printf("c  sigma_b200: clause database built on the device (%u originals, %u learnts)\n", orgs.size(), learnts.size());
		}
		else {
			// receive into a fresh SCNF exactly like shrinkSimp() builds one (simplify.cpp:235-247)
			SCNF fresh;
			fresh.init(sz.n_clauses ? sz.n_clauses : 1, sz.n_literals ? sz.n_literals : 1);
			if (sz.n_buckets) fresh.STYPE::alloc(sz.n_buckets);
			fresh.refs().resize(sz.n_clauses);
			out.arena = sz.n_buckets ? (uint32_t*)fresh.clause(0) : (uint32_t*)st.data();
			out.refs = (uint64_t*)fresh.refs().data();
			out.arena_cap = sz.n_buckets ? sz.n_buckets : 1; out.refs_cap = sz.n_clauses;
			if (!out.refs) out.refs = (uint64_t*)st.data();
			rc = sigma_download(ctx, &out);
			if (rc != SIGMA_OK) { fprintf(stderr, "c sigma_b200: download: %s\n", sigma_last_error(ctx)); exit(EXIT_FAILURE); }
			fresh.migrateTo(scnf);
		}
		for (uint32 v = 1; v <= V; ++v) sp->state[v].state = st[v];
		for (uint32 i = oldtrail; i < trail.size(); ++i) sp->level[ABS(trail[i])] = 0;     // enqueueUnit, solve.hpp:379
		sp->propagated = sz.propagated;
#ifdef STATISTICS
		sigma_report rep;
		sigma_get_report(ctx, &rep);
		BVESTATS& b = stats.simplify.bve;
		b.pures += rep.pures; b.inverters += rep.inverters; b.andors += rep.andors; b.ites += rep.ites; b.xors += rep.xors;
		b.funs += rep.funs; b.resolutions += rep.resolutions;
		stats.simplify.sub.subsumed += rep.subsumed; stats.simplify.sub.strengthened += rep.strengthened;
#endif
	}

	// Solver::simplifying() (simplify.cpp:155-233) with :180-210 replaced
	void sigmaSimplifying() {
		SLEEPING(sleep.simplify, opts.simplify_sleep_en);
		rootify();
		shrinkTop(false);
		if (orgs.empty()) {
			PREFETCH_CM(cs, deleted);
			recycleWT(cs, deleted);
			return;
		}
		timer.stop();
		timer.solve += timer.cpuTime();
		if (!opts.profile_simplifier) timer.start();
		const char* he = getenv("SIGMA_HOST_EXTRACT");
		const bool deviceExtract = !(he && he[0] == '1');
		if (deviceExtract) sigmaAwaken();
		else awaken();
		if (simpstate == AWAKEN_FAIL) { recycle(); return; }
		if (INTERRUPTED) killSolver();
		const int64 bmelted = inf.maxMelted;
		sigmaPass(deviceExtract);
		const int64 bclauses = clausesBefore, bliterals = literalsBefore;
		occurs.clear(true), ot.clear(true);
		const bool success = (bliterals != inf.nLiterals);
		stats.simplify.all.variables += int64(inf.maxMelted) - bmelted;
		stats.simplify.all.clauses += bclauses - int64(inf.nClauses);
		stats.simplify.all.literals += bliterals - int64(inf.nLiterals);
		last.shrink.removed = stats.shrunken;
		if (inf.maxFrozen > sp->simplified) stats.units.forced += inf.maxFrozen - sp->simplified;
		if (!inf.unassigned || !inf.nClauses) { SET_SAT; printStats(1, 's', CGREEN); return; }
		if (newBeginningDone) { /* simplify.cpp:216 happened on the device */ }
		else if (canMap()) map(true);
		else newBeginning();
		rebuildWT(opts.simplify_priorbins);
		if (retrail()) LOG2(2, " Propagation after simplify proved a contradiction");
		UPDATE_SLEEPER(simplify, success);
		printStats(1, 's', CGREEN);
		if (!opts.profile_simplifier) timer.stop(), timer.simplify += timer.cpuTime();
		if (!opts.solve_en) killSolver();
		timer.start();
	}

	// Solver::simplify() (simplify.cpp:102-116)
	void sigmaSimplify() {
		if (!opts.phases && !(opts.all_en || opts.ere_en)) return;
		stats.simplify.calls++;
		sigmaSimplifying();
		INCREASE_LIMIT(simplify, stats.simplify.calls, nlognlogn, true);
		last.simplify.reduces = stats.reduces + 1;
		if (opts.phases > 2) opts.phases--;
	}

	// Solver::solve() (solve.cpp:155-181)
	void sigmaSolve() {
		timer.start();
		initLimits();
		if (verbose == 1) printTable();
		if (canPreSimplify()) sigmaSimplify();
		if (UNSOLVED) {
			MDMInit();
			while (UNSOLVED && !EXHAUSTED) {
				if (BCP()) analyze();
				else if (!inf.unassigned) SET_SAT;
				else if (canReduce()) reduce();
				else if (canRestart()) restart();
				else if (canRephase()) rephase();
				else if (canSimplify()) sigmaSimplify();
				else if (canProbe()) probe();
				else if (canMMD()) MDM();
				else decide();
			}
		}
		timer.stop();
		timer.solve += timer.cpuTime();
		wrapup();
		if (ctx) { sigma_destroy(ctx); ctx = nullptr; }
	}
};

} // namespace

int main(int argc, char** argv)
{   // main.cpp:38-98
	BOOL_OPT opt_competition_en("competition", "engage SAT competition mode", false);
	BOOL_OPT opt_quiet_en("quiet", "enable quiet mode, same as verbose=0", false);
	INT_OPT opt_verbose("verbose", "set the verbosity", 1, INT32R(0, 4));
	INT_OPT opt_timeout("timeout", "set out-of-time limit in seconds", 0, INT32R(0, INT32_MAX));
	INT_OPT opt_memoryout("memoryout", "set out-of-memory limit in gigabytes", 0, INT32R(0, 256));
	if (argc == 1) { fprintf(stderr, "usage: seqfrost_sigma <formula.cnf> [options]\n"); return EXIT_FAILURE; }
	try {
		parseArguments(argc, argv);
		competition_en = opt_competition_en;
		quiet_en = opt_quiet_en, verbose = opt_verbose;
		if (quiet_en || competition_en) verbose = 0;
		else if (!verbose || competition_en) quiet_en = true;
		signal_handler(handler_terminate);
		std::string formula = argv[1];
		SigmaSolver* s = new SigmaSolver(formula);
		solver = s;
		if (opt_timeout > 0) set_timeout(opt_timeout);
		if (opt_memoryout > 0) set_memoryout(opt_memoryout);
		signal_handler(handler_mercy_interrupt, handler_mercy_timeout);
		s->sigmaSolve();
		solver = NULL;
		fflush(stdout);
		_exit(EXIT_SUCCESS);
	}
	catch (std::bad_alloc&) { LOGSAT("UNKNOWN"); return EXIT_FAILURE; }
	catch (MEMOUTEXCEPTION&) { LOGSAT("UNKNOWN"); return EXIT_FAILURE; }
}
