mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_vae_launches.csv python - > gpurun_out/r2_vae_ncu.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from panopticdiffusionmodels_b200.libs.autoencoder import get_model
dev = torch.device("cuda:0")
torch.manual_seed(0)
vae = get_model(None, 0.23010).to(dev)
z = torch.randn(32, 4, 32, 32, device=dev)
vae.decode(z, max_batch=32)
torch.cuda.synchronize()
vae.decode(z, max_batch=32)
torch.cuda.synchronize()
PY
ls -la gpurun_out/r02_vae_launches.csv
