"""Import alias: the product package lives in the directory ``text-to-video-finetuning_b200/`` (not a valid
Python identifier), so ``import t2v_b200`` simply points its package search path there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "text-to-video-finetuning_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
