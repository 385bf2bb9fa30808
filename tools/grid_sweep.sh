for cfg in "8 8 8" "12 8 8" "16 8 8" "8 6 8" "8 12 8" "8 8 4" "8 8 12" "16 6 8"; do
  set -- $cfg
  echo "=== setup=$1 finish=$2 states=$3"
  MLH_GRID_SETUP=$1 MLH_GRID_FINISH=$2 MLH_GRID_STATES=$3 python tools/quick_bench.py sedov61 kh1000j 2>&1 | grep -E "N=|k4a|k4b1|k4b3"
done
