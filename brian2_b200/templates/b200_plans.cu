#ifndef _B200_PLANS_H
#define _B200_PLANS_H
#include "network.h"
{% for plan in plans %}
extern const B200Plan _b200_plan_{{plan.index}};
{% endfor %}
#endif
