/*
 * TEST INFRASTRUCTURE -- pthread stand-in for the slice of boost::thread the
 * reference uses (cmake/CMakeLists.txt:6; src/EAR.cpp:196-207, src/Settings.cpp:56,
 * src/HelperFunctions.cpp:56-75).  Only used to compile the reference into
 * oracle/_ref/.  C++98-compatible because that build uses -std=gnu++98.
 */
#ifndef EAR_B200_BOOST_MUTEX_SHIM
#define EAR_B200_BOOST_MUTEX_SHIM
#include <pthread.h>
namespace boost {
class mutex {
	pthread_mutex_t m;
	mutex(const mutex&);
	mutex& operator=(const mutex&);
public:
	mutex() { pthread_mutex_init(&m, 0); }
	~mutex() { pthread_mutex_destroy(&m); }
	void lock() { pthread_mutex_lock(&m); }
	void unlock() { pthread_mutex_unlock(&m); }
	class scoped_lock {
		mutex& ref;
	public:
		explicit scoped_lock(mutex& mm) : ref(mm) { ref.lock(); }
		~scoped_lock() { ref.unlock(); }
	};
};
}
#endif
