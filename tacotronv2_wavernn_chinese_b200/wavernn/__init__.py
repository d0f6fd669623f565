"""Host-side mirror of the reference's `wavernn` package for the generation path.

Same module names, class names, argument meaning and error behaviour as
lturing/tacotronv2_wavernn_chinese (`wavernn/models/fatchord_version.py`,
`wavernn/utils/{__init__,dsp,paths,display}.py`); the compute goes to libb200tts.so.
"""
