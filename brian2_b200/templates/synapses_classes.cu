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
#include "b200_host.h"
// the reference's generated host code relies on this (it leaks from brianlib/spikequeue.h there)
using namespace std;
typedef b200::Pathway SynapticPathway;
#endif
{% endmacro %}
