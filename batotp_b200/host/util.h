// util.h — the handful of util helpers batest's driver uses (reference: batotp/util.h:46-59,
// util.cpp:52-149).  Same names, argument meaning and return conventions.
#ifndef BATOTP_B200_UTIL_H
#define BATOTP_B200_UTIL_H
#include <cstdint>

struct Time {
  int64_t a;  // seconds
  int64_t b;  // nanoseconds
};
Time getTime(void);                                           // util.cpp:52-70 (CLOCK_REALTIME)
double diffTime(const Time &endTime, const Time &startTime);  // util.cpp:78-88
int doesFileExist(const char *fname);                         // util.cpp:111-125: 0 exists, -1 not
int mkDirIfNec(const char *dirname);                          // util.cpp:133-149: 0 created, -1 otherwise
#endif
