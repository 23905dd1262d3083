// batest — same command line, call sequence, console report and output files as the reference
// driver (test/main.cpp:30-116), running the time-optimisation step on the GPU.
#include <cstdio>
#include <string>
#include <vector>

#include "ba.h"
#include "util.h"

using namespace BATOTP;

int main(int argc, char *argv[]) {
  BA myBA;
  Traj myTraj;
  std::vector<Time> t(6);
  std::string configFileIn;
  if (argc > 1) {  // config file from the terminal argument: everything lives in ./
    configFileIn = argv[1];
    myBA.setHomeFolder("./");
    myBA.setInputFolder("./");
    myBA.setOutputFolder("./");
  } else {
    configFileIn = "config.dat";
    mkDirIfNec(myBA.getOutputFolder().c_str());
  }
  std::string filename = myBA.getInputFolder() + configFileIn;
  myBA.setIsAutoIntegRes(false);

  t[0] = getTime();
  if (myBA.readConfigData(filename.c_str()) == -1) return -1;
  if (myBA.loadTrajectoryData(myTraj) == -1) return -1;
  t[1] = getTime();
  printf("-----Interpolation of input data-----------\n");
  if (myBA.interpInputData(myTraj) == -1) return -1;
  t[2] = getTime();
  printf("\n--Constant-step accel. constraint integ.--\n");
  myBA.setIntegDir(-1);
  myBA.setIsLastSweep(false);
  if (myBA.sweep(myTraj) == -1) return -1;
  myBA.setIntegDir(1);
  myBA.setIsLastSweep(true);
  if (myBA.sweep(myTraj) == -1) return -1;
  t[3] = getTime();
  printf("---------------------------------------\n");
  myBA.interpOutputData(myTraj);
  t[4] = getTime();
  myBA.writeOutputData(myTraj);
  t[5] = getTime();

  printf("\nComputational times (sec):\n");
  printf("Input data interp.      : %f\n", diffTime(t[2], t[1]));
  printf("Accel. constraint integ.: %f\n", diffTime(t[3], t[2]));
  printf("Reading input data      : %f\n", diffTime(t[1], t[0]));
  printf("Writing Output data     : %f\n", diffTime(t[5], t[4]));
  printf("Total,     with file IO : %f\n", diffTime(t[5], t[0]));
  printf("Total,  without file IO : %f\n", diffTime(t[4], t[1]));
  printf("\n");

  filename = myBA.getOutputFolder() + "compTimes.dat";
  FILE *fid = fopen(filename.c_str(), "wb");
  if (fid) {
    float v = (float)diffTime(t[3], t[2]);
    fwrite(&v, 4, 1, fid);
    v = (float)diffTime(t[4], t[1]);
    fwrite(&v, 4, 1, fid);
    v = (float)diffTime(t[5], t[0]);
    fwrite(&v, 4, 1, fid);
    fclose(fid);
  }
  return 0;
}
