#include "sqlite_dyn.h"

#include <dlfcn.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace b2mh {
namespace sq {
namespace {

template <typename F>
void bind(void* lib, const char* name, F& out) {
  void* p = dlsym(lib, name);
  if (!p) throw std::runtime_error(std::string("[sqlite_dyn.cc] libsqlite3 lacks symbol ") + name);
  out = reinterpret_cast<F>(p);
}

Api load() {
  void* lib = nullptr;
  for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
    lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (lib) break;
  }
  if (!lib) throw std::runtime_error("[sqlite_dyn.cc] cannot load libsqlite3.so.0 (needed for COLMAP databases)");
  Api a{};
  bind(lib, "sqlite3_open_v2", a.open_v2);
  bind(lib, "sqlite3_close", a.close);
  bind(lib, "sqlite3_exec", a.exec);
  bind(lib, "sqlite3_free", a.free);
  bind(lib, "sqlite3_prepare_v2", a.prepare_v2);
  bind(lib, "sqlite3_step", a.step);
  bind(lib, "sqlite3_reset", a.reset);
  bind(lib, "sqlite3_clear_bindings", a.clear_bindings);
  bind(lib, "sqlite3_finalize", a.finalize);
  bind(lib, "sqlite3_bind_int64", a.bind_int64);
  bind(lib, "sqlite3_bind_double", a.bind_double);
  bind(lib, "sqlite3_bind_null", a.bind_null);
  bind(lib, "sqlite3_bind_blob64", a.bind_blob64);
  bind(lib, "sqlite3_bind_text", a.bind_text);
  bind(lib, "sqlite3_column_int64", a.column_int64);
  bind(lib, "sqlite3_column_double", a.column_double);
  bind(lib, "sqlite3_column_blob", a.column_blob);
  bind(lib, "sqlite3_column_bytes", a.column_bytes);
  bind(lib, "sqlite3_column_text", a.column_text);
  bind(lib, "sqlite3_column_type", a.column_type);
  bind(lib, "sqlite3_errmsg", a.errmsg);
  bind(lib, "sqlite3_last_insert_rowid", a.last_insert_rowid);
  bind(lib, "sqlite3_libversion", a.libversion);
  return a;
}

}  // namespace

const Api& api() {
  static const Api a = load();  // thread-safe static initialisation; rethrows on every call if it failed
  return a;
}

}  // namespace sq
}  // namespace b2mh
