for NB in 0 1; do
if [ $NB = 1 ]; then export FB_COPY_STREAMS2=1; fi
timeout 200 python bench.py --streams 8 --steps 200 --no-single --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('two_copy_streams=$NB value %.0f (%.1f us) e2e %.0f (%.1f us, link frac %.2f) e2e_sync %.0f (%.1f us)'%(d['value'],1e3*d['ms_per_step'],d['e2e']['value'],1e3*d['e2e']['ms_per_step'],d['e2e']['h2d_link_frac'],d['e2e_sync']['value'],1e3*d['e2e_sync']['ms_per_step']))"; done
python -c "
import torch; p=torch.cuda.get_device_properties(0); print('asyncEngineCount', getattr(p,'async_engine_count', None))"
nvidia-smi -q | grep -i -A3 "pci" | head -30
