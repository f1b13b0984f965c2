"""Column layout of states / observations and of two-sided quantities (reference: gym/index_names.py:1-7)."""
CASH_INDEX, INVENTORY_INDEX, TIME_INDEX, ASSET_PRICE_INDEX = 0, 1, 2, 3
BID_INDEX, ASK_INDEX = 0, 1
