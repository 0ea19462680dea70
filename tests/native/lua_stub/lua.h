/* Minimal DECLARATIONS of the Lua 5.2 C API subset lua-mapreduce_b200/lua/mrhbm_lua.c uses -- test infrastructure:
 * this image has no Lua, so the binding cannot be built or run here; with these prototypes (written from the Lua 5.2
 * reference manual, section 4.8) `gcc -fsyntax-only -Wall` at least type-checks every call it makes.
 * Not a Lua implementation; never linked. */
#ifndef LUA_STUB_H
#define LUA_STUB_H
#include <stddef.h>
typedef struct lua_State lua_State;
typedef double lua_Number;
typedef ptrdiff_t lua_Integer;
typedef int (*lua_CFunction)(lua_State *L);
#define LUA_TNUMBER 3
#define LUA_TTABLE 5
#define LUA_REGISTRYINDEX (-1001000)
#define lua_upvalueindex(i) (LUA_REGISTRYINDEX - (i))
void lua_pushnil(lua_State *L);
const char *lua_pushstring(lua_State *L, const char *s);
const char *lua_pushlstring(lua_State *L, const char *s, size_t len);
void lua_pushboolean(lua_State *L, int b);
void lua_pushnumber(lua_State *L, lua_Number n);
void lua_pushinteger(lua_State *L, lua_Integer n);
void lua_pushvalue(lua_State *L, int idx);
void lua_pushcclosure(lua_State *L, lua_CFunction fn, int n);
#define lua_pushcfunction(L, f) lua_pushcclosure(L, (f), 0)
void lua_getfield(lua_State *L, int idx, const char *k);
void lua_setfield(lua_State *L, int idx, const char *k);
void lua_rawseti(lua_State *L, int idx, int n);
void lua_createtable(lua_State *L, int narr, int nrec);
void *lua_newuserdata(lua_State *L, size_t sz);
void lua_setuservalue(lua_State *L, int idx);
void *lua_touserdata(lua_State *L, int idx);
int lua_toboolean(lua_State *L, int idx);
lua_Number lua_tonumberx(lua_State *L, int idx, int *isnum);
#define lua_tonumber(L, i) lua_tonumberx(L, (i), NULL)
int lua_type(lua_State *L, int idx);
void lua_settop(lua_State *L, int idx);
#define lua_pop(L, n) lua_settop(L, -(n)-1)
#endif
