"""B200-native text-to-video finetune hot path (drop-in for ExponentialML/Text-To-Video-Finetuning's
``models`` / ``utils.lora*`` / ``train.py`` surface).  Host code is Python/PyTorch; every kernel on the path is
hand-written sm_100a CUDA behind the C ABI in ``include/t2v_b200.h`` (see ``native.py``)."""
__version__ = "0.1.0"
