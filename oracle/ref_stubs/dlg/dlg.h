/* TEST INFRASTRUCTURE: stand-in for <dlg/dlg.h> (dlg v0.3): the four macros /root/reference/include/logging.h:13-16
 * maps LOG() to, printing "[level] message" lines to stderr (see ../README.md). */
#pragma once
#include <stdio.h>
#define REF_DLG(level, ...) do { fprintf(stderr, "[" level "] "); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } while (0)
#define dlg_debug(...) do { if (0) fprintf(stderr, __VA_ARGS__); } while (0)
#define dlg_info(...) REF_DLG("info", __VA_ARGS__)
#define dlg_warn(...) REF_DLG("warn", __VA_ARGS__)
#define dlg_error(...) REF_DLG("error", __VA_ARGS__)
