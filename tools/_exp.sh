timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "variants or nerf_rgb or train_gradients or create_nerf" 2>&1 | tail -25
