{# USES_VARIABLES { _n_sources, _n_targets, delay, _source_dt} #}
{# "Pushing" spikes costs nothing on this device: the thresholder has already written the spike
   list into the current slot of the source's spike ring (see threshold.cu), and delays are
   resolved at delivery time by the delay-binned CSR.  What remains of
   brian2/devices/cpp_standalone/templates/synapses_push_spikes.cpp is the before_run block,
   which builds that CSR (replacing SynapticPathway::prepare / CSpikeQueue::prepare,
   synapses_classes.cpp:63-86, spikequeue.h:48-105). #}
{% extends 'common_group.cu' %}

{% block b200_file %}
// ===== code object {{codeobj_name}}: no device work (spike ring makes push_spikes a no-op) =====
void _run_{{codeobj_name}}() {}
{% endblock %}

{% block before_code %}
    {% set scalar = c_data_type(variables['delay'].dtype) %}
    std::vector<{{scalar}}> &real_delays = {{get_array_name(variables['delay'], access_data=False)}};
    {{scalar}}* real_delays_data = real_delays.empty() ? 0 : &(real_delays[0]);
    std::vector<int32_t> &_b200_srcs = {{get_array_name(owner.synapse_sources, access_data=False)}};
    std::vector<int32_t> &_b200_tgts = {{get_array_name(owner.synapse_targets, access_data=False)}};
    const size_t n_delays = real_delays.size();
    const size_t n_synapses = _b200_srcs.size();
    {{owner.name}}.prepare({{b200_host_constant_or_scalar('_n_sources', variables['_n_sources'])}},
                           {{b200_host_constant_or_scalar('_n_targets', variables['_n_targets'])}},
                           real_delays_data, n_delays,
                           _b200_srcs.empty() ? 0 : &_b200_srcs[0],
                           _b200_tgts.empty() ? 0 : &_b200_tgts[0],
                           n_synapses, {{_source_dt}},
                           &_b200_es{{get_array_name(eventspace_variable, access_data=False)}},
                           {{'true' if owner.prepost == 'post' else 'false'}}, (int64_t){{b200_post_parent_size}});
{% endblock %}
