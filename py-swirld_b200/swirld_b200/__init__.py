"""swirld_b200 -- B200-native virtual-voting engine for py-swirld's consensus
hot path (can_see / divide_rounds / decide_fame / find_order)."""
