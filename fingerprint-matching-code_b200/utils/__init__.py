"""``utils`` of the B200 matching head: the hot-path modules of ``/root/reference/utils`` (``hungarian``,
``feature_align``, ``factorize_graph_matching``, ``pad_tensor``, ``data_to_cuda``, ``build_graphs``).

The reference's ``utils`` is a regular package that also holds host-side helpers this repo does not rebuild
(``models_sl``, ``scheduler``, ``visualize``, ``matching``, ``augmentation``), and its scripts import both kinds
(``train.py:14-26``, ``evaluate_binary_classifier.py:27-34``).  So this package OVERLAYS the reference's instead of
shadowing it: every other ``utils`` directory on ``sys.path`` (i.e. the reference root, listed after this one) is
appended to ``__path__``, and a module missing here resolves there.  ``src`` needs no such code: the reference's ``src``
and ``src/model`` have no ``__init__.py``, and neither have ours, so they are namespace packages that span both roots
with this repo's modules first.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
