"""Host-side mirror of the reference's Python interface for the capture-zone path.

Same module/function names, argument meaning and error behaviour as
oneka/{model,probabilityfield,capturezone,stochastic,deterministic,utilities}.py of the
reference; the hot loops are gone -- they are enqueued on the GPU through the C ABI.
"""
