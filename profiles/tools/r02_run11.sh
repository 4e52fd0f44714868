for v in "A" "B NMC_OVERLAP_TARGETS=0" "C NMC_SIREN_TC_BWD=0" "D NMC_OVERLAP_TARGETS=0 NMC_SIREN_TC_BWD=0" "E NMC_NO_GRAPH=1 NMC_OVERLAP_TARGETS=0 NMC_SIREN_TC_BWD=0"; do
  set -- $v; name=$1; shift
  echo "=== variant $name: $@"
  env "$@" timeout 600 python profiles/tools/taylor_green_run.py 4 shipped 2>&1 | grep -E "^initial|^step" | cut -c 1-110
done
