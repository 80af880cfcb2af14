"""Import shim: lets the reference's `import nvdiffrast.torch as dr` (mesh_sfs_optim.py:14, neural_render.py,
train_mlp.py, train_unet.py, get_data.py) resolve to the fmhr_b200 operators when this repo root is on sys.path."""
__version__ = "0.3.1+fmhr_b200"
