#!/bin/bash
# static SASS instruction counts per kernel (and opcode mix for one kernel): tools/sass_count.sh [kernel-regex]
cd "$(dirname "$0")/.."
OBJ=sequential_monte_carlo_b200/lib/smcb_filter.o
[ -n "$2" ] && OBJ=$2
cuobjdump -sass $OBJ | awk -v pat="${1:-.}" '
/Function :/ {name=$3; next}
/^\s+\/\*[0-9a-f]+\*\/\s+[A-Z@]/ { if (name ~ pat) { n[name]++; op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;/,"",op); c[name" "op]++ } }
END { for (k in n) print n[k], k; }' | sort -n | sed 's/_ZN4smcb[0-9]*_GLOBAL__N__[0-9a-f_]*smcb_[a-z]*_cu_[0-9a-f]*//' 
