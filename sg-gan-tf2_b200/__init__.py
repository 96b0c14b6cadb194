"""sg-gan-tf2_b200 -- B200-native SG-GAN training step behind the reference's Python surface.

The directory name is not a Python identifier; load it with
    importlib.import_module("sg-gan-tf2_b200")
or put the directory on sys.path and `import model, module, ops` exactly like the reference's
flat script layout (main.py:10 `from model import sggan`).
"""
