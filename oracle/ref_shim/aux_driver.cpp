// aux_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN read_inputfile<T>()
// (src/acquisition.cxx:31-59) on stdin and prints what it returns, so the product's AuxData text
// reader can be pinned to it (tests/golden/aux_parser.json).  The function body is not copied into
// this repository: oracle/Makefile extracts those lines from /root/reference into the git-ignored
// oracle/_ref/read_inputfile.inc at build time and this file includes the result.
// usage: ref_read_inputfile float|double < text
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "read_inputfile.inc"

int main(int argc, char **argv)
{
    const bool as_float = argc > 1 && !std::strcmp(argv[1], "float");
    if (as_float) {
        for (float v : read_inputfile<float>(&std::cin)) std::printf("%.9g\n", (double)v);
    } else {
        for (double v : read_inputfile<double>(&std::cin)) std::printf("%.17g\n", v);
    }
    return 0;
}
