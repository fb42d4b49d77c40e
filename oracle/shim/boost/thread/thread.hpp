/*
 * TEST INFRASTRUCTURE -- pthread stand-in for boost::thread / boost::thread_group
 * (see mutex.hpp).  create_thread copies the functor, as boost does, so the copy
 * shares the Recorder* of the original (src/SceneContext.h:28,46-48).
 * Each worker folds its thread-local gmtl shim counters into a process-wide total
 * when it ends, so the harness can report segment counts for threaded renders.
 */
#ifndef EAR_B200_BOOST_THREAD_SHIM
#define EAR_B200_BOOST_THREAD_SHIM
#include <pthread.h>
#include <vector>
#include "mutex.hpp"
#include <gmtl/gmtl.h>
namespace boost {
struct shim_totals {
	static unsigned long long& ray_tests() { static unsigned long long v = 0; return v; }
	static unsigned long long& seg_tests() { static unsigned long long v = 0; return v; }
	static mutex& lock() { static mutex m; return m; }
	static void fold() {
		mutex::scoped_lock l(lock());
		gmtl::ShimCounters& c = gmtl::shim_counters();
		ray_tests() += c.ray_tests; seg_tests() += c.seg_tests;
		c.ray_tests = 0; c.seg_tests = 0;
	}
};
class thread {
public:
	pthread_t handle;
	template <class F> static void* trampoline(void* p) {
		F* f = static_cast<F*>(p);
		(*f)();
		delete f;
		shim_totals::fold();
		return 0;
	}
	template <class F> explicit thread(F f) {
		F* copy = new F(f);
		pthread_create(&handle, 0, &thread::trampoline<F>, copy);
	}
	void join() { pthread_join(handle, 0); }
};
class thread_group {
	std::vector<thread*> threads;
public:
	~thread_group() { for (size_t i = 0; i < threads.size(); ++i) delete threads[i]; }
	template <class F> thread* create_thread(F f) {
		thread* t = new thread(f);
		threads.push_back(t);
		return t;
	}
	void join_all() { for (size_t i = 0; i < threads.size(); ++i) threads[i]->join(); }
};
}
#endif
