/* TEST TOOLING ONLY: LD_PRELOAD shim that pins time() so that `ema align -d`, which seeds libc rand()
 * from time() (src/split.c:54-58 in the reference), is reproducible.  Both the reference binary and the
 * ema-b200 CLI are run under it by tests/test_gpu_platforms.py. */
#include <time.h>
time_t time(time_t *t)
{
	const time_t v = 1234567890;
	if (t) *t = v;
	return v;
}
