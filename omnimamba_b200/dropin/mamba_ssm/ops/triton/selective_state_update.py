from omnimamba_b200.interface.selective_state_update import selective_state_update  # noqa: F401
