/*
 * TEST INFRASTRUCTURE -- linked into oracle/_ref/EAR_ref only.  The reference seeds
 * rand() from time() (src/EAR.cpp:58, src/Scene.cpp:116); overriding time() with an
 * environment-provided value makes `EAR_ref calc T60` reproducible so its printed
 * T60 can be compared with the oracle for the same seed.  Unset => real clock.
 */
#include <stdlib.h>
#include <time.h>
#include <sys/time.h>
extern "C" time_t time(time_t* out) {
	const char* s = getenv("EAR_REF_SEED");
	time_t v;
	if (s) v = (time_t)atol(s);
	else { struct timeval tv; gettimeofday(&tv, 0); v = tv.tv_sec; }
	if (out) *out = v;
	return v;
}
