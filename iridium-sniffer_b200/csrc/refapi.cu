// refapi.cu -- placeholder, filled in below (reference-shaped per-item API + plug-in ABI).
