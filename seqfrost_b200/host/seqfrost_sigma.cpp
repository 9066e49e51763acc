// seqfrost_sigma.cpp - the `seqfrost` command line with its SIGmA pass running on the B200: the reference's main()
// (main.cpp:38-98) over SigmaSolver (sigma_solver.hpp).  Usage is the reference CLI's:
//
//     seqfrost_sigma <formula.cnf> [reference options]
//
// Build: g++ -std=c++17 -O3 -DNDEBUG -DSTATISTICS -I<reference>/src
//     seqfrost_sigma.cpp libseqfrost.a -L. -lsigma_b200
#include "sigma_solver.hpp"

using namespace SeqFROST;

bool quiet_en = false;
bool competition_en = false;
int verbose = -1;

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
