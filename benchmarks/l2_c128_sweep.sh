# complex128 sweep path (n = 20, 16 MiB per chain): L2-resident group schedule knobs.
for cfg in "0 1" "16 4" "16 3" "16 2" "32 2" "16 5" "48 1"; do
  set -- $cfg
  echo "== c128 MIB=$1 STREAMS=$2"
  QUMCMC_L2_GROUP_MIB=$1 QUMCMC_L2_STREAMS=$2 python benchmarks/configs.py c128 | cut -c1-260
done
