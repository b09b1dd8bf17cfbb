"""Batched counterparts of the reference's ``optimax_rogue/game`` package."""
