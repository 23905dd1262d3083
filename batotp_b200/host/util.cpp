#include "util.h"

#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

Time getTime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  Time t;
  t.a = (int64_t)ts.tv_sec;
  t.b = (int64_t)ts.tv_nsec;
  return t;
}

double diffTime(const Time &e, const Time &s) {
  if (e.b < s.b) return (double)(e.a - s.a) - 1 + (double)(e.b - s.b + 1.0e9) / 1.0e9;
  return (double)(e.a - s.a) + (double)(e.b - s.b) / 1.0e9;
}

int doesFileExist(const char *fname) { return access(fname, F_OK) != -1 ? 0 : -1; }

int mkDirIfNec(const char *dirname) {
  return mkdir(dirname, S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) == -1 ? -1 : 0;
}
