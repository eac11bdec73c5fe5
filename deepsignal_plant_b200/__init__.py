"""deepsignal_plant_b200 -- B200-native (sm_100a) implementation of deepsignal-plant's
per-site methylation classifier hot path, behind the reference's own Python API.

Public surface (mirrors ``deepsignal_plant``):

* ``models.ModelBiLSTM``              -- drop-in for ``deepsignal_plant.models.ModelBiLSTM``
* ``call_modifications._call_mods``   -- drop-in for the batch step that drives it
* ``call_mods_freq``                  -- per-site frequency aggregation (call_freq)
* ``freq_dist``                       -- the same across the GPUs of one box (torchrun; records exchanged over NVLink)
* ``chain``                           -- call_mods -> call_freq without text in between (record columns on the device)

All arithmetic runs in hand-written CUDA kernels inside ``libdsp_b200.so`` (C ABI in
``include/dsp_b200.h``); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
