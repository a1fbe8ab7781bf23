{# The reference's SynapticPathway/CSpikeQueue pair (templates/synapses_classes.cpp:15-88) is
   replaced by b200::Pathway (csrc/b200_host.h): a delay-binned CSR on the device plus the
   source's spike ring. #}
{% macro cpp_file() %}
{% endmacro %}
{% macro h_file() %}
#ifndef _BRIAN_SYNAPSES_H
#define _BRIAN_SYNAPSES_H
#include <vector>
#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include "b200_host.h"
// the reference's generated host code relies on this (it leaks from brianlib/spikequeue.h there)
using namespace std;
typedef b200::Pathway SynapticPathway;
// The reference's generated connect() code reports invalid indices / sample sizes with
// `cout << "Error: ..."; exit(1);` (synapses_create_generator.cpp:103-106,194-197) -- fine for a
// ./main process, fatal for a library living inside the Python process.  Inside this project it
// becomes an exception that the C ABI turns into a non-zero status ("Project run failed").
#include <cstdlib>
#include <stdexcept>
namespace b200 {
[[noreturn]] inline void exit_from_generated_code(int code) {
    throw std::runtime_error("generated host code called exit(" + std::to_string(code) +
                             ") (see the error message printed above)");
}
}
#define exit(code) b200::exit_from_generated_code(code)
#endif
{% endmacro %}
