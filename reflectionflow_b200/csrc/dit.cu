// placeholder until the DiT orchestrator lands
