"""Market-model descriptors (midprice, arrivals, fills, price impact) flattened into mbt_config."""
