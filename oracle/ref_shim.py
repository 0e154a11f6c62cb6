"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* reference.

Only usable in the build container, where `/root/reference` is mounted
(it does not exist on the GPU box).  Used by `oracle/make_golden.py` to
generate the committed fixtures under `tests/golden/` and by the
`not gpu` tests that cross-check the oracle port against the real
reference when it happens to be present.

The reference cannot be imported as-is: `train/__init__.py` pulls in
`torchmetrics` (evaluate.py:12-13) and `train/utils.py:3` imports
`matplotlib.pyplot`; neither is installed and neither touches the loss
arithmetic, so both are replaced by empty stub modules.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('USL_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'train', 'loss.py'))


def import_reference():
    """Returns (loss_module, utils_module, sparsification_module)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')

    sys.dont_write_bytecode = True  # the tree is read-only

    for name in ('matplotlib', 'matplotlib.pyplot',
                 'torchmetrics', 'torchmetrics.functional'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    functional = sys.modules['torchmetrics.functional']
    if not hasattr(functional, 'structural_similarity_index_measure'):
        functional.structural_similarity_index_measure = None

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import warnings
    warnings.filterwarnings('ignore', message='.*align_corners.*')

    import train.loss as ref_loss
    import train.utils as ref_utils
    import train.sparsification as ref_spars
    return ref_loss, ref_utils, ref_spars
