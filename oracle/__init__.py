"""CPU oracle for the Parler-TTS generate() hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch fp32/bf16 + numpy)
restatement of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the *checker* or the
timed CPU baseline -- never as a product path.  ``parler_tts_b200`` must not
import anything from here.

Parity pinning status:
  * delay pattern, ParlerTTSLogitsProcessor, decoder forward: PINNED against the
    reference's own code executed in the build container (fixtures under
    ``tests/golden/`` written by ``tests/golden/make_golden.py``).
  * ``_sample`` glue: restated from transformers 4.46.1 semantics (the loop is
    not in /root/reference and not runnable under transformers 5.5) -- unpinned.
  * DAC decode: pinned against ``transformers.models.dac.DacModel`` (an
    independent restatement of descript-audio-codec, which is not installed) --
    "parity unpinned" against the reference's real dependency.
"""
