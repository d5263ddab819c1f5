// Package banner (reference src/version.hpp:19-27 prints MLSL_PACKAGE_VERSION at Init).
#pragma once
#define MLSLB_PRODUCT_VERSION "2026.1"
#define MLSLB_PACKAGE_VERSION "mlsl-b200 " MLSLB_PRODUCT_VERSION " (Blackwell sm_100a, API-compatible with Intel(R) MLSL 2018 API 1.0)"
