// sigma_solver.hpp - the reference-side binding of libsigma_b200: a SeqFROST::Solver whose SIGmA pass runs on the B200.
//
// Everything except the replaced region is the reference's own code, linked UNMODIFIED from libseqfrost(.a): parser,
// CDCL search, options, the incremental API, model extension/verification, the SCLAUSE hand-off (newBeginning ->
// addClause(SCLAUSE&), sclause.cpp:22-62) and the watch rebuild.  Solver::solve(), isolve(), simplify() and simplifying()
// are not virtual, so this subclass re-states their few lines of top-level control flow (solve.cpp:155-181,
// solveinc.cpp:80-116, simplify.cpp:102-116,155-233) and swaps the phase loop simplify.cpp:180-210 for the C-ABI calls of
// include/sigma_b200.h.  By default extract() (simplify.cpp:118-133) and newBeginning() (simplify.cpp:249-268) run on the
// device as well; SIGMA_HOST_EXTRACT=1 / SIGMA_HOST_NEWBEGINNING=1 keep the reference's own.
//
// Used by seqfrost_sigma.cpp (the `seqfrost` CLI) and by embedding programs through sigmaSolve() / sigmaISolve()
// exactly where they would call solve() / isolve() (tests/incr/incr_driver.cpp).
#ifndef SIGMA_SOLVER_HPP
#define SIGMA_SOLVER_HPP

#include "control.hpp"
#include "solve.hpp"
#include "can.hpp"
#include "version.hpp"
#include "../../include/sigma_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace SeqFROST {

// The device context is created on a helper thread that starts BEFORE the reference's constructor runs (a base class is
// constructed first): CUDA initialisation and loading the kernels take about a second, the parser next to it several.
struct SigmaCtxStarter {
	sigma_ctx* early = nullptr;
	int earlyrc = SIGMA_OK;
	std::thread starter;
	SigmaCtxStarter() {
		const char* lazy = getenv("SIGMA_LAZY_CREATE");
		if (lazy && lazy[0] == '1') return;
		starter = std::thread([this]() {
			const char* dev = getenv("SIGMA_DEVICE");
			earlyrc = sigma_create(dev ? atoi(dev) : 0, &early);
		});
	}
	~SigmaCtxStarter() { if (starter.joinable()) starter.join(); if (early) sigma_destroy(early); }
	// hands the context over (once); nullptr when it was not started here
	sigma_ctx* take(int& rc) {
		if (starter.joinable()) starter.join();
		sigma_ctx* c = early;
		early = nullptr;
		rc = earlyrc;
		return c;
	}
};

class SigmaSolver : private SigmaCtxStarter, public Solver {
	sigma_ctx* ctx = nullptr;
	std::vector<uint8_t> st;
	int64 clausesBefore = 0, literalsBefore = 0;      // inf.nClauses / inf.nLiterals after extract()
	bool newBeginningDone = false;                    // cm / orgs / learnts already hold the simplified formula
	bool watchesDone = false;                         // ... and wt holds what rebuildWT(opts.simplify_priorbins) would build
	std::vector<uint32_t> wtoffs;                     // sigma_download_watches: list offsets and the WATCH records
	struct WatchRec { uint64_t ref; uint32_t imp; int32_t size; };
	std::vector<WatchRec> wtrecs;

	// SIGMA_BINDING_TIMES=1: wall-clock split of one simplifying() call on stderr (where the host side of the drop-in goes)
	std::chrono::steady_clock::time_point tmark;
	bool timesOn = false;
	void tstart() { const char* e = getenv("SIGMA_BINDING_TIMES"); timesOn = e && e[0] == '1'; tmark = std::chrono::steady_clock::now(); }
	void tsplit(const char* what) {
		if (!timesOn) return;
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "c sigma_b200 time: %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tmark).count());
		tmark = now;
	}

	// Options the device pass does not cover are refused once, before any solving, instead of ending a solve half way
	// (sigma_execute would answer SIGMA_E_UNSUPPORTED for them).
	void checkSupportedOptions() const {
		const char* why = nullptr;
		if (opts.proof_en) why = "-proof (DRAT lines from inside the pass)";
		else if (opts.ere_en && opts.ere_clause_max > 256) why = "--redundancyclausemax above 256";
		else if (opts.ere_en && opts.ere_max_occurs > 65535) why = "--redundancymaxoccurs above 65535";
		if (why) { fprintf(stderr, "c sigma_b200: option not supported by the device pass: %s\n", why); exit(EXIT_FAILURE); }
	}

	[[noreturn]] void sigmaFail(const int rc, const char* where) {
		fprintf(stderr, "c sigma_b200: %s: %s: %s\n", where, sigma_strerror(rc), ctx ? sigma_last_error(ctx) : "");
		if (rc == SIGMA_E_NOMEM) throw MEMOUTEXCEPTION();     // the reference's own out-of-memory path (malloc.hpp:28-58, main.cpp:93)
		exit(EXIT_FAILURE);                                   // no CPU fallback, by design
	}

public:
	explicit SigmaSolver(const std::string& path) : Solver(path) { checkSupportedOptions(); }
	SigmaSolver() : Solver() { checkSupportedOptions(); }     // incremental constructor (solveinc.cpp:26-49)
	~SigmaSolver() { if (wtFreer.joinable()) wtFreer.join(); if (ctx) { sigma_destroy(ctx); ctx = nullptr; } }

	// awaken() (simplify.cpp:135-153) without extract(): the SCNF and the occurrence table live on the device
	// wt.clear(true) (simplify.cpp:140) frees millions of small watch vectors - more than half a second on the 10M-clause
	// instance.  The table is handed to a helper thread that frees it while the copies and the pass run; the solver's `wt` is
	// empty from here on, exactly as after the reference's call.
	WT oldwt;
	std::thread wtFreer;
	void sigmaAwaken() {
		printStats(1, '-', CGREEN0);
		initSimp();
		if (wtFreer.joinable()) wtFreer.join();
		wt.migrateTo(oldwt);                               // wt: no memory, size 0 == wt.clear(true)
		wtFreer = std::thread([this]() { oldwt.clear(true); });
		inf.nClauses = inf.nLiterals = 0;
	}

	// simplify.cpp:180-210 on the GPU.  Inputs and outputs are the reference's own objects.
	void sigmaPass(const bool deviceExtract) {
		if (!ctx) {
			int rc = SIGMA_OK;
			ctx = take(rc);                                    // started next to the parser
			if (!ctx && rc == SIGMA_OK) {
				const char* dev = getenv("SIGMA_DEVICE");
				rc = sigma_create(dev ? atoi(dev) : 0, &ctx);
			}
			if (rc != SIGMA_OK) sigmaFail(rc, "sigma_create");
			// INTERRUPTED polls (simplify.cpp:176,206; subsume.cpp:26; bounded.cpp:31): the pass reads the solver's own flag,
			// which Solver::interrupt() sets from the signal handlers (control.cpp:137-149)
			sigma_set_interrupt_flag(ctx, (const volatile unsigned char*)&interrupted);
			tsplit("sigma_create");
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
			// the clause arena is malloc'ed (pageable)
			const char* hr = getenv("SIGMA_HOST_REGISTER");
			// (SIGMA_HOST_REGISTER=1: pin it for the copy.  Off by default: registering and unregistering 400 MB costs about a second,
			// the copy from pageable memory 30 ms)
			const bool pin = hr && hr[0] == '1' && db.cm_buckets >= (1u << 20);
			const bool pinned = pin && sigma_host_register((void*)db.cm, db.cm_buckets * sizeof(uint32_t)) == SIGMA_OK;
			rc = sigma_upload_cnf(ctx, &o, &db, &in);
			if (pinned) sigma_host_unregister((void*)db.cm);
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
		tsplit("upload (+ device extract)");
		if (rc == SIGMA_OK) rc = sigma_execute(ctx);
		tsplit("sigma_execute");
		if (rc == SIGMA_UNSAT) { learnEmpty(); killSolver(); }           // propelim.cpp:60, subsume.cpp:66, bounded.cpp:74
		if (rc == SIGMA_INTERRUPTED) killSolver();                        // simplify.cpp:206, subsume.cpp:26, bounded.cpp:31
		if (rc != SIGMA_OK) sigmaFail(rc, "pass");
		sigma_formula sz = {};
		sigma_result_sizes(ctx, &sz);
		inf.nClauses = sz.n_clauses; inf.nLiterals = (uint32)sz.n_literals;
		inf.unassigned = sz.unassigned; inf.maxFrozen = sz.max_frozen; inf.maxMelted = sz.max_melted;
		// the solver state first (state / value / root trail / model stack): the SAT-by-elimination check, the statistics and
		// vmap.initiate() of map(true) read it before any clause comes back
		const uint32 oldtrail = trail.size();
		trail.resize(sz.trail_size);
		model.resolved.resize((uint32)sz.model_size);
		sigma_formula out = {};
		out.state = st.data(); out.value = (int8_t*)sp->value; out.trail = trail.data(); out.model = model.resolved.data();
		out.trail_cap = sz.trail_size; out.model_cap = sz.model_size;
		if (!out.trail) out.trail = (uint32_t*)st.data();
		if (!out.model) out.model = (uint32_t*)st.data();
		rc = sigma_download_state(ctx, &out);
		if (rc != SIGMA_OK) sigmaFail(rc, "sigma_download_state");
		for (uint32 v = 1; v <= V; ++v) sp->state[v].state = st[v];
		for (uint32 i = oldtrail; i < trail.size(); ++i) sp->level[ABS(trail[i])] = 0;     // enqueueUnit, solve.hpp:379
		sp->propagated = sz.propagated;
		tsplit("sigma_download_state");
		sigma_report rep;
		sigma_get_report(ctx, &rep);
		if (opts.profile_simplifier) {
			// -profilesimplifier (statistics.cpp:34-42): the reference accumulates host milliseconds per stage
			// (timer.pstart/pstop around createOT, the ordering, sortOT, SUB, VE, ERE, shrinkSimp); here the same fields
			// receive the CUDA-event milliseconds of the corresponding kernel groups
			timer.cot += rep.stage_ms[SIGMA_T_OT_HIST] + rep.stage_ms[SIGMA_T_OT_SCAN] + rep.stage_ms[SIGMA_T_OT_SCATTER] +
			             rep.stage_ms[SIGMA_T_OT_ORDER] + rep.stage_ms[SIGMA_T_OT_UPDATE];
			timer.vo += rep.stage_ms[SIGMA_T_VO];
			timer.sot += rep.stage_ms[SIGMA_T_SOT];
			timer.sub += rep.stage_ms[SIGMA_T_SUB];
			timer.ve += rep.stage_ms[SIGMA_T_VE] + rep.stage_ms[SIGMA_T_NBH];
			timer.ere += rep.stage_ms[SIGMA_T_ERE];
			timer.gc += rep.stage_ms[SIGMA_T_GC] + rep.stage_ms[SIGMA_T_EXPORT];   // in ms (the reference adds seconds here, simplify.cpp:246)
		}
#ifdef STATISTICS
		BVESTATS& b = stats.simplify.bve;
		b.pures += rep.pures; b.inverters += rep.inverters; b.andors += rep.andors; b.ites += rep.ites; b.xors += rep.xors;
		b.funs += rep.funs; b.resolutions += rep.resolutions;
		stats.simplify.sub.subsumed += rep.subsumed; stats.simplify.sub.strengthened += rep.strengthened;
#endif
	}

	// simplify.cpp:215-218: `if (canMap()) map(true); else newBeginning();` and rebuildWT() - from the device when possible
	// (newBeginningDone / watchesDone say what was done here; the caller runs the reference's own code for the rest)
	void sigmaHandBack() {
		const uint32 V = inf.maxVar;
		int rc;
		sigma_formula sz = {};
		sigma_result_sizes(ctx, &sz);
		const char* hn = getenv("SIGMA_HOST_NEWBEGINNING");
		newBeginningDone = false;
		watchesDone = false;
		if (!(hn && hn[0] == '1') && inf.unassigned && inf.nClauses && !opts.aggr_cnf_sort) {
			// map(true), map.cpp:24-46: the mapping is computed on the host exactly as the reference does, the clauses come back
			// in the new numbering (mapLit on every literal, sclause.cpp:34-35 / map.hpp:195-205)
			const bool mapping = canMap();
			if (mapping) {
				assert(!LEVEL);
				stats.mapping.calls++;
				vmap.initiate(sp);
				vmap.mapOrgs(model.lits);
				if (assumptions.size()) vmap.mapOrgs(assumptions);
				if (ifrozen.size()) vmap.mapShrinkVars(ifrozen);
				vmap.mapTransitive(last.transitive.literals);
				rc = sigma_set_var_map(ctx, *vmap, vmap.numVars());
				if (rc != SIGMA_OK) sigmaFail(rc, "sigma_set_var_map");
				tsplit("map(true): vmap");
			}
			sigma_cnf_out co = {};
			rc = sigma_cnf_out_sizes(ctx, &co);
			if (rc != SIGMA_OK) sigmaFail(rc, "sigma_cnf_out_sizes");
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
			const char* hr = getenv("SIGMA_HOST_REGISTER");
			const bool pin = hr && hr[0] == '1' && co.cm_buckets >= (1u << 20);
			const bool pinned = pin && sigma_host_register(co.cm, co.cm_buckets * sizeof(uint32_t)) == SIGMA_OK;
			// rebuildWT(opts.simplify_priorbins) (simplify.cpp:218) on the device as well: it runs sortClause on the CLAUSE records,
			// so it comes before they are copied back
			const char* hw = getenv("SIGMA_HOST_REBUILDWT");
			sigma_watches ws = {};
			tsplit("cm.init + bulk alloc");
			const bool deviceWT = !(hw && hw[0] == '1') && sigma_watches_sizes(ctx, opts.simplify_priorbins, &ws) == SIGMA_OK;
			tsplit("device newBeginning + rebuildWT");
			sigma_formula none = {};                                           // clause database only: the state came back before
			rc = sigma_download_cnf(ctx, &co, &none);
			tsplit("sigma_download_cnf");
			if (pinned) sigma_host_unregister(co.cm);
			if (rc != SIGMA_OK) sigmaFail(rc, "sigma_download_cnf");
			if (deviceWT) {
				wtoffs.resize(ws.nlit + 1);
				wtrecs.resize(ws.n_records ? ws.n_records : 1);
				ws.offs = wtoffs.data(); ws.offs_cap = wtoffs.size();
				ws.records = wtrecs.data(); ws.records_cap = wtrecs.size();
				rc = sigma_download_watches(ctx, &ws);
				if (rc != SIGMA_OK) sigmaFail(rc, "sigma_download_watches");
				tsplit("sigma_download_watches");
				// the lists arrive ready-made, in push order; Vec<WATCH> by Vec<WATCH> (solvetypes.hpp:31-32), a few host threads
				// side by side (one allocation per list).  ws.nlit == inf.nDualVars, after map(true) the new one
				if (wtFreer.joinable()) wtFreer.join();
				wt.resize(ws.nlit);
				const uint32 nl = ws.nlit;
				auto fill = [this](const uint32 lo, const uint32 hi) {
					for (uint32 lit = lo; lit < hi; ++lit) {
						const uint32 n = wtoffs[lit + 1] - wtoffs[lit];
						if (!n) continue;
						WL& list = wt[lit];
						list.resize(n);
						memcpy((WATCH*)list, &wtrecs[wtoffs[lit]], size_t(n) * sizeof(WATCH));
					}
				};
				const uint32 nthreads = nl >= (1u << 18) ? 4 : 1;
				std::vector<std::thread> fillers;
				for (uint32 t = 1; t < nthreads; ++t) fillers.emplace_back(fill, uint32(uint64_t(nl) * t / nthreads), uint32(uint64_t(nl) * (t + 1) / nthreads));
				fill(0, uint32(uint64_t(nl) / nthreads));
				for (auto& th : fillers) th.join();
				watchesDone = true;
				tsplit("watch lists into wt");
			}
			for (uint32 v = 1; v <= V; ++v) if (sub[v]) sp->state[v].subsume = 1;   // mark_ssubsume, sclause.hpp:209-215 (old numbering)
			stats.literals.original = co.lits_original; stats.literals.learnt = co.lits_learnt;
			stats.clauses.original = orgs.size(); stats.clauses.learnt = learnts.size();
			scnf.destroy();
			newBeginningDone = true;
			if (verbose >= 2) printf("c  sigma_b200: clause database built on the device (%u originals, %u learnts%s)\n", orgs.size(), learnts.size(), mapping ? ", mapped" : "");
			if (mapping) sigmaMapTail();
		}
		else {
			// receive into a fresh SCNF exactly like shrinkSimp() builds one (simplify.cpp:235-247); map(true) / newBeginning() and
			// rebuildWT() are then the reference's own
			sigma_formula out = {};
			out.state = st.data(); out.value = (int8_t*)sp->value; out.trail = trail.size() ? trail.data() : (uint32_t*)st.data();
			out.model = model.resolved.size() ? model.resolved.data() : (uint32_t*)st.data();
			out.trail_cap = sz.trail_size; out.model_cap = sz.model_size;
			SCNF fresh;
			fresh.init(sz.n_clauses ? sz.n_clauses : 1, sz.n_literals ? sz.n_literals : 1);
			if (sz.n_buckets) fresh.STYPE::alloc(sz.n_buckets);
			fresh.refs().resize(sz.n_clauses);
			out.arena = sz.n_buckets ? (uint32_t*)fresh.clause(0) : (uint32_t*)st.data();
			out.refs = (uint64_t*)fresh.refs().data();
			out.arena_cap = sz.n_buckets ? sz.n_buckets : 1; out.refs_cap = sz.n_clauses;
			if (!out.refs) out.refs = (uint64_t*)st.data();
			rc = sigma_download(ctx, &out);
			if (rc != SIGMA_OK) sigmaFail(rc, "sigma_download");
			fresh.migrateTo(scnf);
		}
	}

	// the rest of Solver::map(true) after its newBeginning() (map.cpp:47-121), unchanged
	void sigmaMapTail() {
		// map trail, queue and heap
		vmap.mapShrinkLits(trail);
		sp->propagated = trail.size();
		const uint32 firstFrozen = vmap.firstL0();
		// - VMTF
		vmtf.map(*vmap, firstFrozen, vmap.size());
		vmap.mapShrinkVars(bumps);
		// - VSIDS
		uVec1D tmp;
		while (vsidsheap.size()) {
			uint32 x = vsidsheap.pop();
			if (x == firstFrozen) continue;
			uint32 mx = vmap.mapped(x);
			if (mx) tmp.push(mx);
		}
		vmap.mapShrinkVars(vsids.scores);
		vsidsheap.rebuild(tmp);
		// - CHB
		tmp.clear();
		while (chbheap.size()) {
			uint32 x = chbheap.pop();
			if (x == firstFrozen) continue;
			uint32 mx = vmap.mapped(x);
			if (mx) tmp.push(mx);
		}
		vmap.mapShrinkVars(chb.scores);
		vmap.mapShrinkVars(chb.conflicts);
		chbheap.rebuild(tmp);
		// shrink others without mapping
		eligible.clear(true);
		eligible.reserve(vmap.numVars());
		elected.clear(true);
		elected.reserve(vmap.numVars());
		occurs.clear(true);
		dlevel.clear(true);
		dlevel.reserve(vmap.numVars());
		dlevel.push(level_t());
		// map search space
		SP* newSP = new SP(vmap.size(), opts.polarity);
		vmap.mapSP(newSP);
		delete sp;
		sp = newSP;
		// map model and proof variables
		vmap.mapShrinkVars(vorg);
		model.init(vorg);
		if (opts.proof_en) proof.init(sp, vorg);
		// update phase-saving counters
		sp->trailpivot = 0, last.rephase.best = last.rephase.target = 0;
		for (uint32 v = 1; v <= vmap.numVars(); ++v) {
			if (sp->state[v].state) continue;
			Phase_t& ph = sp->phase[v];
			if (!PHASE_UNASSIGNED(ph.best)) last.rephase.best++;
			if (!PHASE_UNASSIGNED(ph.target)) last.rephase.target++;
		}
		stats.mapping.compressed += inf.maxVar - vmap.numVars();
		LOG2(2, " Variable mapping compressed %d to %d", inf.maxVar, vmap.numVars());
		inf.maxVar = vmap.numVars();
		inf.nDualVars = V2L(inf.maxVar + 1);
		inf.maxFrozen = inf.maxMelted = inf.maxSubstituted = 0;
		vmap.destroy();
		printStats(1, 'a', CGREEN2);
		tsplit("map(true): search state");
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
		tstart();
		if (deviceExtract) sigmaAwaken();
		else awaken();
		tsplit("awaken (host part)");
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
		sigmaHandBack();
		if (newBeginningDone) { /* simplify.cpp:215-216 happened on the device (map(true) included) */ }
		else if (canMap()) map(true);
		else newBeginning();
		tsplit("state / statistics");
		if (wtFreer.joinable()) wtFreer.join();
		tsplit("wait for the old watch table to be freed");
		if (!(newBeginningDone && watchesDone)) rebuildWT(opts.simplify_priorbins);
		tsplit("host rebuildWT (if any)");
		if (retrail()) LOG2(2, " Propagation after simplify proved a contradiction");
		tsplit("retrail");
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

	// Solver::isolve() (solveinc.cpp:80-116): incremental solving under assumptions.  LCVE on the device skips the
	// variables the embedding program froze (ifreeze) or assumed (iassume): their `ifrozen` bytes travel with the upload.
	void sigmaISolve(Lits_t& assumptions) {
		timer.start();
		iallocSpace();
		iunassume();
		if (BCP()) learnEmpty();
		else {
			initLimits();
			iassume(assumptions);
			if (verbose == 1) printTable();
			if (canPreSimplify()) sigmaSimplify();
			if (UNSOLVED) {
				MDMInit();
				while (UNSOLVED && !INTERRUPTED) {
					if (BCP()) analyze();
					else if (!inf.unassigned) SET_SAT;
					else if (canReduce()) reduce();
					else if (canRestart()) restart();
					else if (canRephase()) rephase();
					else if (canSimplify()) sigmaSimplify();
					else if (canProbe()) probe();
					else if (canMMD()) MDM();
					else idecide();
				}
			}
		}
		timer.stop();
		timer.solve += timer.cpuTime();
		wrapup();
	}

	// what the embedding program reads after (i)solve: verdict and the search statistics (tests compare them)
	int verdict() const { return SAT ? 10 : (UNSAT ? 20 : 0); }
	int64 conflictsSoFar() const { return (int64)stats.conflicts; }
};


} // namespace SeqFROST

#endif
