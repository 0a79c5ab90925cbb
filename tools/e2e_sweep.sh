# usage: bash tools/e2e_sweep.sh -- sweeps the item-group / lane configuration of bgp_dense_elbo_grad_host (e2e line of bench.py)
for cfg in "4 12" "4 6" "2 4" "2 6" "3 9"; do
  set -- $cfg
  echo "conc=$1 lanes=$2: $(BGP_HOST_CONC=$1 BGP_HOST_LANES=$2 python bench.py --steps 2 --warmup 1 --no-cpu --no-fit 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"])')"
done
