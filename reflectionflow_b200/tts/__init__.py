"""Outer search loop of ReflectionFlow (reference: tts/): noise scaling, reflection rounds,
verifier hooks, candidate sharding over the GPUs of one node."""
